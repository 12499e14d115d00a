"""Generates tests/golden/golden.json from the CPU oracle (run in the authoring container).

The reference holds no golden bitstreams and cannot be built here (no Rust), so these vectors pin the
ORACLE against regressions and give the GPU tests fixed targets; they are not outputs of the real `orz` binary.
Inputs are seeded generators from tests/datagen.py plus, when data/enwik8.xz is present, a prefix of enwik8.
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402
from tests import datagen  # noqa: E402

cases = {
    "empty": b"",
    "a": b"a",
    "abcabc": b"abcabcabcabcabcabcabcabc",
    "fox": b"the quick brown fox jumps over the lazy dog. " * 20,
    "text_64k": datagen.text_like(65536, seed=101),
    "random_20k": datagen.random_bytes(20000, seed=102),
    "lowent_100k": datagen.low_entropy(100000, seed=103),
}
e8 = datagen.enwik8(2_000_000)
out = {"_about": "oracle outputs (tools/make_golden.py); sha256 of the orz stream per level; full hex for tiny inputs", "cases": {}}
for name, data in cases.items():
    entry = {"len": len(data), "sha256_in": hashlib.sha256(data).hexdigest(), "levels": {}}
    for level in (0, 1, 2):
        enc = pyoracle.encode(data, level)
        assert pyoracle.decode(enc, len(data) + 64) == data
        entry["levels"][str(level)] = {"len": len(enc), "sha256": hashlib.sha256(enc).hexdigest()}
        if len(enc) <= 256:
            entry["levels"][str(level)]["hex"] = enc.hex()
    out["cases"][name] = entry
if e8 is not None:
    entry = {"len": len(e8), "sha256_in": hashlib.sha256(e8).hexdigest(), "levels": {}}
    for level in (0, 1, 2):
        enc = pyoracle.encode(e8, level)
        entry["levels"][str(level)] = {"len": len(enc), "sha256": hashlib.sha256(enc).hexdigest()}
    out["enwik8_2M"] = entry
# whole-file sizes measured with the oracle next to the README's published sizes (README.md:43-46)
out["enwik8_full_sizes"] = {"oracle": {"0": 27863045, "1": 27182097, "2": 26856813},
                            "readme": {"0": 27898433, "1": 27217825, "2": 26892825},
                            "note": "oracle is 0.13% smaller at every level; the README predates commit 29414c6 (crate 1.6.1 vs tag v1.6.2)"}
os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "golden.json"), "w"), indent=1)
print("wrote", len(out["cases"]), "cases")
