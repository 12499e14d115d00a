"""Repeated one-stream encodes with per-step provenance (A/B build with -DORZ_V7_PROV=1, ORZ_B200_PROV=1): when a stream differs
from the oracle's, the first differing parse step is located (tools/diff_parse.py) and the provenance of the records around it is
printed.  `python tools/prov_probe.py NBYTES LEVEL REPS [KEY=VALUE...]`.  Development tool."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    n, level, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    os.environ["ORZ_B200_PROV"] = "1"
    for a in sys.argv[4:]:
        k, v = a.split("=", 1)
        os.environ[k] = v
    from tools import textgen, diff_parse
    from orz_b200 import Batch
    from orz_b200._lib import load
    from oracle import pyoracle
    data, _ = textgen.corpus(n)
    data = data[:n].tobytes()
    want = pyoracle.encode(data, level)
    L = load()
    L.orz_b200_debug_prov.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32)]
    bad = 0
    for rep in range(reps):
        b = Batch(Batch.ENCODE, [len(data)], cfg=level, device=0)
        b.upload(0, data)
        b.run()
        got = b.download(0)
        if got != want:
            bad += 1
            d = diff_parse.first_difference(got, want)
            print(f"== rep {rep}: differs, len {len(got)} vs {len(want)}", flush=True)
            if d is not None:
                diff_parse.show(d, 4)
                ci, k, before, ra, rb = d
                X = int(rb[k][0])  # input position of the first differing step
                for y in range(X - 1, X + 4):
                    w = (C.c_uint32 * 16)()
                    rc = L.orz_b200_debug_prov(b.h, y, w)
                    F, cnt, rmn, gc, nci, s3, e0x, e0y, T, how, r3, q1, x1, x2, bcr, s1 = list(w)
                    print(f"  prov pos {y} rc {rc}: build F {F} cnt {cnt} r {rmn & 255} mode {(rmn >> 8) & 255} n0 {(rmn >> 16) & 255} n1 {rmn >> 24} gc {gc} | ncand {nci & 255} incomplete {(nci >> 8) & 255} "
                          f"best {(nci >> 16) & 255} mlen {s3 & 255} lazy1 {(s3 >> 8) & 255} lazy2 {(s3 >> 16) & 255} exp {s3 >> 24} | newest entry pos {e0x & 0x1ffffff} ord {e0y & 0x7fffffff} "
                          f"| last repair: T {T} how {how & 255} seen {(how >> 8) & 1} rmask {(how >> 12) & 7} nbr {how >> 16} mlen {r3 & 255} lazy1 {(r3 >> 8) & 255} lazy2 {(r3 >> 16) & 255} q1 {q1} "
                          f"| the repair used: key {x1 & 0x1fffff} ncand {(x1 >> 21) & 63} F_ext {x2} k {bcr & 255} reached {(bcr >> 8) & 1} bc {bcr >> 12} nv0 {s1 & 255} dmax {(s1 >> 8) & 255} nold {(s1 >> 16) & 255} lcp0 {s1 >> 24}", flush=True)
        b.close()
    print(f"== {bad} of {reps} repetitions differ", flush=True)


main()
