"""Runs one-stream encodes until one differs from the oracle's stream and prints the input position of the first differing parse
step: `python tools/find_diff_pos.py NBYTES LEVEL MAXREPS` -> last line `POS <input position>` (or `POS none`).  Development tool."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    n, level, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    from tools import textgen, diff_parse
    from orz_b200 import Batch
    from oracle import pyoracle
    data, _ = textgen.corpus(n)
    data = data[:n].tobytes()
    want = pyoracle.encode(data, level)
    for rep in range(reps):
        b = Batch(Batch.ENCODE, [len(data)], cfg=level, device=0)
        b.upload(0, data)
        b.run()
        got = b.download(0)
        b.close()
        if got != want:
            d = diff_parse.first_difference(got, want)
            diff_parse.show(d, 5)
            print(f"rep {rep} differs")
            print(f"POS {int(d[4][d[1]][0])}")
            return
    print("POS none")


main()
