"""One-shot GPU probe of the opt-in stage A v4 (ORZ_B200_PARSE=4) against v3: parity with the oracle and parse time."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orz_b200 import Batch
from oracle import pyoracle
from tests import datagen

def run(data, level, env):
    for k in ("ORZ_B200_PARSE", "ORZ_B200_PARSE_WARPS", "ORZ_B200_STATS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    b = Batch(Batch.ENCODE, [len(data)], cfg=level)
    b.upload(0, data)
    t0 = time.time()
    try:
        b.run()
    except Exception as e:  # noqa: BLE001
        return {"env": env, "error": repr(e)}
    dt = time.time() - t0
    out = b.download(0)
    ms, k = b.timing(0)
    b.close()
    return {"env": env, "n": len(data), "wall_s": round(dt, 4), "parse_ms": round(ms, 2), "launches": k, "out_len": len(out),
            "mb_s_parse": round(len(data) / ms / 1e3, 3) if ms else None, "out": out}

def main():
    res = []
    small = datagen.text_like(300_000, seed=5)
    e8 = datagen.enwik8(4_000_000) or datagen.text_like(4_000_000, seed=6)
    want_small = {l: pyoracle.encode(small, l) for l in (0, 2)}
    want_e8 = pyoracle.encode(e8, 2)
    for env in ({}, {"ORZ_B200_PARSE": "4"}, {"ORZ_B200_PARSE": "4", "ORZ_B200_PARSE_WARPS": "16"}):
        for l in (0, 2):
            r = run(small, l, env)
            r["identical"] = r.pop("out", None) == want_small[l]
            r["case"] = f"text300k_l{l}"
            res.append(r)
            print(r, flush=True)
        r = run(e8, 2, env)
        r["identical"] = r.pop("out", None) == want_e8
        r["case"] = "enwik8_4M_l2"
        res.append(r)
        print(r, flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/v4_probe.json", "w"), indent=1)

main()
