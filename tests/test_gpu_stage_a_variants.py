"""GPU parity of every stage A kernel in the tree.  v3 and v4 were compared with the oracle on a B200 in round 1; the
v5 / v6 families (parse_v5.cuh, parse_v6.cuh) were written after the round's GPU budget was spent and have only run in
the SIMT emulator, so their tests are expected-to-fail-or-pass (xfail, non-strict): the round-end run on a GPU box
records which of them are bit-identical on hardware without turning the suite red.  Each variant runs in its own
process (tests/gpu_variant_case.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case(*switches, timeout=300):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "gpu_variant_case.py"), *switches], capture_output=True, text=True,
                         timeout=timeout)
    assert out.returncode == 0 and "identical" in out.stdout, (out.stdout[-500:], out.stderr[-1500:])


@pytest.mark.parametrize("switches", [("ORZ_B200_PARSE=3",), ("ORZ_B200_PARSE=4", "ORZ_B200_PARSE_WARPS=16"),
                                      ("ORZ_B200_PARSE=4", "ORZ_B200_PARSE_WARPS=32")])
def test_measured_kernels(switches):
    _case(*switches)


EXPERIMENTAL = [
    ("ORZ_B200_PARSE=5", "ORZ_B200_PARSE_WARPS=32"),
    ("ORZ_B200_PARSE=5", "ORZ_B200_PARSE_WARPS=16"),
    ("ORZ_B200_PARSE=5", "ORZ_B200_PARSE_CLUSTER=2"),
    ("ORZ_B200_PARSE=5", "ORZ_B200_PARSE_CLUSTER=8"),
    ("ORZ_B200_PARSE=6", "ORZ_B200_PARSE_GROUP=8", "ORZ_B200_PARSE_WARPS=32"),
    ("ORZ_B200_PARSE=6", "ORZ_B200_PARSE_GROUP=16", "ORZ_B200_PARSE_WARPS=16"),
    ("ORZ_B200_PARSE=6", "ORZ_B200_PARSE_GROUP=4", "ORZ_B200_PARSE_WARPS=16"),
    ("ORZ_B200_PARSE=6", "ORZ_B200_PARSE_GROUP=8", "ORZ_B200_PARSE_CLUSTER=4"),
]


@pytest.mark.xfail(strict=False, reason="written after the round-1 GPU budget was spent: emulator-verified only")
@pytest.mark.parametrize("switches", EXPERIMENTAL)
def test_experimental_kernels(switches):
    _case(*switches, timeout=75)  # (a normal run takes ~15 s; the kernels' own watchdog gives up after a few seconds without progress)
