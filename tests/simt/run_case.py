"""Runs inside a subprocess of tests/test_simt_emulation.py with ORZ_B200_LIB pointing at the emulator build.

Encodes the requested seeded inputs through the ordinary Python host mirror (C ABI -> kernels, here executed by the
SIMT emulator) and prints one JSON line: per case the sha256 and length of the produced stream.  The parent test
compares them with the oracle.  TEST INFRASTRUCTURE ONLY.
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import orz_b200  # noqa: E402
from tests import datagen  # noqa: E402


def main():
    spec = json.loads(sys.argv[1])
    datas = []
    for c in spec["cases"]:
        if "literal" in c:
            datas.append(bytes.fromhex(c["literal"]))
        else:
            datas.append(getattr(datagen, c["gen"])(c["n"], seed=c.get("seed", 1)))
    if spec.get("operators"):  # LZEncoder::{new,encode,forward} driven chunk by chunk from the host (lz.rs:75-95)
        outs = [orz_b200.encode_with_operators(d, spec["level"]) for d in datas]
    elif spec.get("batch"):
        outs = orz_b200.encode_batch(datas, spec["level"])
    else:
        outs = [orz_b200.encode(d, spec["level"]) for d in datas]
    print(json.dumps([{"len": len(o), "sha256": hashlib.sha256(o).hexdigest()} for o in outs]))


if __name__ == "__main__":
    main()
