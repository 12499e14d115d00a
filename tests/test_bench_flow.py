"""CPU tier: bench.py's own control flow (autotune adoption, first-pass oracle check, fallback, JSON keys) executed
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
               ORZ_BENCH_NO_VARIANTS="1", SIMT_SWITCH_LIMIT="20000000000")
    env.pop("ORZ_B200_PARSE", None)
    out = subprocess.run([sys.executable, os.path.join(SIMT, "run_bench_flow.py"), mode], env=env, capture_output=True, text=True,
                         timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("mode,kernel", [("default", "orz_parse4_kernel"), ("tuned", "orz_parse6_kernel"),
                                         ("mismatch", "orz_parse4_kernel")])
def test_bench_flow(mode, kernel):
    r = _flow(mode)
    assert r["identical"] is True and r["kernel"] == kernel and REQUIRED <= set(r["keys"])
    if mode == "mismatch":
        assert "default kept" in r["stage_a"]["selection"]
    if mode == "tuned":
        assert r["stage_a"]["switches"]["ORZ_B200_PARSE"] == "6"
