"""Times ONE stage A variant on a slice of the bench corpus and checks its streams against stage A v3's.

Run by bench.py in a subprocess per variant (so that a faulting or hanging experimental kernel cannot take the bench
down): `python tools/variant_probe.py NBYTES NSTREAMS KEY=VALUE ...` — NSTREAMS independent streams of NBYTES each
(consecutive slices of the corpus); the KEY=VALUE pairs are the ORZ_B200_PARSE* switches of the variant.  Prints one
JSON line {"env", "streams", "parse_ms", "mb_s", "identical_to_v3"}: v3 (ORZ_B200_PARSE=3) is the kernel whose
streams were compared with the oracle's over the whole corpus and in every GPU test of round 1.  No oracle involved:
bench.py itself compares the stream of the kernel it ends up timing with the oracle's.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

KEYS = ("ORZ_B200_PARSE", "ORZ_B200_PARSE_WARPS", "ORZ_B200_PARSE_CLUSTER", "ORZ_B200_PARSE_PARK", "ORZ_B200_PARSE_BUILDERS",
        "ORZ_B200_PARSE_GROUP", "ORZ_B200_DBG")


def run(segs, env):
    from orz_b200 import Batch
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(env)
    b = Batch(Batch.ENCODE, [len(s) for s in segs], cfg=2, device=int(os.environ.get("ORZ_PROBE_DEVICE", "0")))
    for i, s in enumerate(segs):
        b.upload(i, s)
    b.run()
    outs = [b.download(i) for i in range(len(segs))]
    ms, _ = b.timing(0)
    b.close()
    return outs, ms


def main():
    n, ns = int(sys.argv[1]), int(sys.argv[2])
    env = dict(a.split("=", 1) for a in sys.argv[3:])
    assert all(k in KEYS for k in env), env
    from tools import textgen
    data, _ = textgen.corpus(n * ns)
    segs = [data[i * n:(i + 1) * n].tobytes() for i in range(ns)]
    want, _ = run(segs, {"ORZ_B200_PARSE": "3"})
    run(segs, env)  # warm-up (module load, attribute set)
    got, ms = run(segs, env)
    print(json.dumps({"env": env, "bytes": n, "streams": ns, "parse_ms": round(ms, 2),
                      "mb_s": round(n * ns / ms / 1e3, 3) if ms else None, "identical_to_v3": got == want}))


if __name__ == "__main__":
    main()
