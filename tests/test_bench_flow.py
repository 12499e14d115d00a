"""CPU tier: bench.py's own control flow (warm-up / timed passes, oracle comparison of the timed stream, JSON keys) executed
against the SIMT-emulator build with a stand-in for torch — the bench line is what the round is judged on, so its
branches are not left to be run for the first time on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIMT = os.path.join(ROOT, "tests", "simt")
REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"}


def _flow(mode):
    subprocess.check_call(["make", "-C", SIMT, "liborz_b200_emu.so"], stdout=subprocess.DEVNULL)
    env = dict(os.environ, ORZ_B200_LIB=os.path.join(SIMT, "liborz_b200_emu.so"), ORZ_B200_TEST_EMULATOR="1",
               ORZ_BENCH_NO_DECODE="1", ORZ_B200_V7_CLUSTER="1", SIMT_SWITCH_LIMIT="20000000000")
    out = subprocess.run([sys.executable, os.path.join(SIMT, "run_bench_flow.py"), mode], env=env, capture_output=True, text=True,
                         timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("mode", ["default", "mismatch"])
def test_bench_flow(mode):
    r = _flow(mode)
    assert REQUIRED <= set(r["keys"]) and r["kernel"].startswith("orz_parse7_kernel")
    assert r["config"]["stage_a"] == r["kernel"] and "workload" in r["config"]
    if mode == "default":
        assert r["identical"] is True and r["config"]["bit_identical_to_oracle"] is True and r["invalid"] == "debug corpus size"
    else:  # a stream that differs from the oracle's must be visible where the driver keeps it, and invalidate the line
        assert r["identical"] is False and r["config"]["bit_identical_to_oracle"] is False and "differs" in r["invalid"]
