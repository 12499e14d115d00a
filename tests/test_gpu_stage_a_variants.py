"""GPU parity of every stage A kernel in the library against the oracle (chunk cut, block slide + forward, deep chains).
Everything that ships is strict: stage A v7 (parse_v7.cuh, the default) with each cluster size the library can launch, and the
ring-record kernels (v3 two block sizes, v4) that serve many-stream batches and LZCfg depths beyond 48.  Each variant runs in
its own process (tests/gpu_variant_case.py) because the library reads its switches once per process."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case(*switches, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "gpu_variant_case.py"), *switches], capture_output=True, text=True,
                         timeout=timeout)
    assert out.returncode == 0 and "identical" in out.stdout, (out.stdout[-500:], out.stderr[-1500:])


@pytest.mark.parametrize("switches", [(), ("ORZ_B200_V7_CLUSTER=1",), ("ORZ_B200_STATS=1",)])
def test_stage_a_v7(switches):
    _case(*switches)


@pytest.mark.parametrize("switches", [("ORZ_B200_LEGACY_STAGE_A=v3",), ("ORZ_B200_LEGACY_STAGE_A=v3s",), ("ORZ_B200_LEGACY_STAGE_A=v4",)])
def test_ring_record_kernels(switches):
    _case(*switches)
