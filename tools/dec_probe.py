"""GPU probe of the decoder: decodes the oracle's stream of a slice of the bench corpus (one stream), checks the bytes,
prints CUDA-event time of the decode kernel.  `python tools/dec_probe.py NBYTES [LEVEL] [KEY=VALUE ...]`.  Development tool."""
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
    comp = pyoracle.encode(data, level)
    res = {}
    for rep in range(2):
        b = Batch(Batch.DECODE, [len(comp)], out_caps=[len(data) + 64], device=0)
        b.upload(0, comp)
        t0 = time.time()
        b.run()
        wall = time.time() - t0
        got = b.download(0)
        ms = b.timing(3)[0]
        res = {"bytes": n, "level": level, "wall_s": round(wall, 4), "decode_ms": round(ms, 2), "mb_s": round(n / ms / 1e3, 3), "identical": got == data}
    print(json.dumps(res))


main()
