"""Repeats the one-stream encode of the bench corpus and counts streams that differ from the oracle's (timing-dependent protocol
bugs show up as rare mismatches).  `python tools/stress_probe.py NBYTES LEVEL REPS`.  Development tool."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    n, level, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    from tools import textgen
    from orz_b200 import Batch
    from oracle import pyoracle
    data, _ = textgen.corpus(n)
    data = data[:n].tobytes()
    want = pyoracle.encode(data, level)
    bad = []
    for rep in range(reps):
        b = Batch(Batch.ENCODE, [len(data)], cfg=level, device=0)
        b.upload(0, data)
        b.run()
        got = b.download(0)
        b.close()
        if got != want:
            m = min(len(got), len(want))
            fd = next((i for i in range(0, m, 4096) if got[i:i + 4096] != want[i:i + 4096]), m)
            fd = next((i for i in range(fd, min(fd + 4096, m)) if got[i] != want[i]), fd)
            bad.append({"rep": rep, "len": len(got), "first_diff": fd})
            out = os.path.join(ROOT, "gpurun_out", f"bad_stream_{len(got)}_{fd}.bin")
            if len(bad) <= 1 and not os.path.exists(out):  # (keep the merge-back small: one sample)
                os.makedirs(os.path.dirname(out), exist_ok=True)
                open(out, "wb").write(got)
    print(json.dumps({"bytes": n, "level": level, "reps": reps, "want_len": len(want), "mismatches": bad}), flush=True)


main()
