"""Subprocess body of tests/test_bench_flow.py: runs bench.main() against the SIMT-emulator build with a stand-in for the
few torch calls bench.py makes (tests/simt/faketorch), so that the bench's control flow — autotune selection, first-pass
check against the oracle, fallback to the default, JSON assembly — is exercised without a GPU.  TEST INFRASTRUCTURE ONLY;
the numbers it prints mean nothing."""
import contextlib
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests", "simt", "faketorch"))
sys.path.insert(0, ROOT)
mode = sys.argv[1]
sys.argv = ["bench.py", "--bytes", "3000", "--steps", "1", "--warmup", "1"]
import bench  # noqa: E402
from oracle import pyoracle  # noqa: E402

V6 = {"ORZ_B200_PARSE": "6", "ORZ_B200_PARSE_GROUP": "8", "ORZ_B200_PARSE_WARPS": "16"}
if mode == "default":
    bench.autotune_stage_a = lambda device=0: ({}, [])
else:
    bench.autotune_stage_a = lambda device=0: (dict(V6), [])
if mode == "mismatch":  # the tuned kernel's first-pass check sees a different stream -> the default must come back
    real, calls = pyoracle.encode, {"n": 0}

    def fake(d, l):
        calls["n"] += 1
        out = real(d, l)
        return out[:-1] + b"\x55" if calls["n"] == 1 else out
    pyoracle.encode = fake
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    bench.main()
o = json.loads(buf.getvalue().strip().splitlines()[-1])
print(json.dumps({"stage_a": o["config"]["stage_a"], "identical": o["bit_identical_to_oracle"], "kernel": o["roofline"]["kernel"],
                  "keys": sorted(o.keys())}))
