"""GPU probe of stage A v7: encodes slices of the bench corpus as one stream, checks the stream against the oracle's,
prints CUDA-event time per kernel class.  `python tools/v7_probe.py NBYTES [LEVEL] [KEY=VALUE ...]` (library switches
such as ORZ_B200_STATS=1, ORZ_B200_LEGACY_STAGE_A=1 go into the environment BEFORE the library
loads).  Development tool, not a bench line."""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    n = int(sys.argv[1])
    level = int(sys.argv[2]) if len(sys.argv) > 2 and "=" not in sys.argv[2] else 2
    for a in sys.argv[2:]:
        if "=" in a:
            k, v = a.split("=", 1)
            os.environ[k] = v
    from tools import textgen
    from orz_b200 import Batch
    from oracle import pyoracle
    data, _ = textgen.corpus(n)
    data = data[:n].tobytes()
    t0 = time.time()
    want = pyoracle.encode(data, level)
    t_cpu = time.time() - t0
    res = {}
    for rep in range(2):
        b = Batch(Batch.ENCODE, [len(data)], cfg=level, device=0)
        b.upload(0, data)
        t0 = time.time()
        b.run()
        wall = time.time() - t0
        got = b.download(0)
        res = {"bytes": n, "level": level, "wall_s": round(wall, 4), "mb_s": round(n / wall / 1e6, 3), "identical": got == want,
               "oracle_mb_s": round(n / t_cpu / 1e6, 2), "out_len": len(got), "want_len": len(want),
               "ms": {name: [round(x, 2) for x in b.timing(c)] for c, name in enumerate(["parse", "emit", "block", "decode", "run"])}}
        if got != want:
            m = min(len(got), len(want))
            fd = next((i for i in range(0, m, 4096) if got[i:i + 4096] != want[i:i + 4096]), m)
            fd = next((i for i in range(fd, min(fd + 4096, m)) if got[i] != want[i]), fd)
            res["first_diff"] = fd
        b.close()
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
