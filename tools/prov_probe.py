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
                for j in range(max(0, k - 3), min(len(ra), k + 4)):
                    w = (C.c_uint32 * 8)()
                    rc = L.orz_b200_debug_prov(b.h, before + j, w)
                    x, s1, s3, s10, F, gc, ord0, hz = list(w)
                    print(f"  prov item {j} rc {rc}: x {x} key {s1 & 0x1fffff} ncand {(s1 >> 21) & 63} incomplete {(s1 >> 27) & 1} | mlen {s3 & 255} lazy1 {(s3 >> 8) & 255} "
                          f"lazy2 {(s3 >> 16) & 255} exp {s3 >> 24} | cnt16 {s10 & 0xffff} r {(s10 >> 16) & 15} n0 {(s10 >> 20) & 63} mode {(s10 >> 26) & 3} "
                          f"last repair {(s10 >> 28) & 3} repairs {s10 >> 30} | F {F} gc {gc} ord0 {ord0} | ty {hz & 255} mlen {(hz >> 8) & 255} lane {(hz >> 16) & 255} nst {hz >> 24}", flush=True)
        b.close()
    print(f"== {bad} of {reps} repetitions differ", flush=True)


main()
