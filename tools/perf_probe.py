"""Quick GPU perf probe (not the bench): encode/decode enwik8 slices, check vs oracle, print kernel timings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import orz_b200
from orz_b200 import Batch
from oracle import pyoracle
from tests import datagen

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
    nstreams = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    level = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    check = (sys.argv[4] != "nocheck") if len(sys.argv) > 4 else True
    full = datagen.enwik8(min(n * nstreams, 100_000_000))
    if full is None:
        full = datagen.text_like(min(n * nstreams, 50_000_000))
    nseg = max(1, len(full) // n)
    segs = [full[(i % nseg) * n:((i % nseg) + 1) * n] for i in range(nstreams)]  # tiles the corpus when it is too short
    b = Batch(Batch.ENCODE, [len(s) for s in segs], cfg=level)
    for i, s in enumerate(segs):
        b.upload(i, s)
    t0 = time.time(); b.run(); t1 = time.time()
    outs = [b.download(i) for i in range(nstreams)]
    tot = sum(len(s) for s in segs)
    print(f"encode l{level}: {nstreams} x {n} B: {t1-t0:.3f} s  {tot/(t1-t0)/1e6:.2f} MB/s  ratio {sum(map(len,outs))/tot:.4f}")
    for c, name in enumerate(["parse", "emit", "block", "decode"]):
        ms, k = b.timing(c)
        if k: print(f"   {name}: {ms:.1f} ms over {k} launches")
    if check:
        t0 = time.time(); want = pyoracle.encode(segs[0], level); t1 = time.time()
        print(f"   oracle (1 core) stream 0: {len(segs[0])/(t1-t0)/1e6:.2f} MB/s; identical: {want == outs[0]}")
    b.close()
    d = Batch(Batch.DECODE, [len(o) for o in outs], out_caps=[len(s) + 64 for s in segs])
    for i, o in enumerate(outs):
        d.upload(i, o)
    t0 = time.time(); d.run(); t1 = time.time()
    ok = all(d.download(i) == segs[i] for i in range(nstreams))
    ms, k = d.timing(3)
    print(f"decode: {t1-t0:.3f} s  {tot/(t1-t0)/1e6:.2f} MB/s  kernel {ms:.1f} ms  roundtrip ok: {ok}")
    if check:
        t0 = time.time(); pyoracle.decode(outs[0], len(segs[0]) + 64); t1 = time.time()
        print(f"   oracle decode (1 core) stream 0: {len(segs[0])/(t1-t0)/1e6:.2f} MB/s")
    d.close()

main()
