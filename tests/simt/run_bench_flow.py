"""Subprocess body of tests/test_bench_flow.py: runs bench.main() against the SIMT-emulator build with a stand-in for the
few torch calls bench.py makes (tests/simt/faketorch), so that the bench's control flow — warm-up / timed passes, the oracle
comparison of the timed stream, JSON assembly — is exercised without a GPU.  TEST INFRASTRUCTURE ONLY; the numbers it prints
mean nothing."""
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

if mode == "mismatch":  # the oracle "disagrees": the line must say so and be marked invalid
    real = pyoracle.encode
    pyoracle.encode = lambda d, l=2, cfg=None: real(d, l)[:-1] + b"\x55"
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    bench.main()
o = json.loads(buf.getvalue().strip().splitlines()[-1])
print(json.dumps({"config": o["config"], "identical": o["bit_identical_to_oracle"], "kernel": o["roofline"]["kernel"],
                  "invalid": o.get("invalid"), "keys": sorted(o.keys())}))
