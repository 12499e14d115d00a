"""Event trace of one window position across repeated one-stream encodes (instrumented stage A): prints, per repetition, whether
the stream matched the oracle's; the library prints the trace (ORZ_B200_STATS=1, ORZ_B200_DBG_POS, ORZ_B200_DBG_INPOS) to stderr.
`python tools/trace_probe.py NBYTES LEVEL REPS KEY=VALUE...`.  Development tool."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    n, level, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    for a in sys.argv[4:]:
        k, v = a.split("=", 1)
        os.environ[k] = v
    from tools import textgen
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
        sys.stderr.flush()
        if got != want:
            out = os.path.join(ROOT, "gpurun_out", f"bad_stream_rep{rep}_{len(got)}.bin")
            if os.environ.get("ORZ_TRACE_SAVE") and sum(f.startswith("bad_stream_rep") for f in os.listdir(os.path.join(ROOT, "gpurun_out"))) < 1:
                open(out, "wb").write(got)
        print(f"== rep {rep} identical {got == want} len {len(got)} want {len(want)}", file=sys.stderr, flush=True)


main()
