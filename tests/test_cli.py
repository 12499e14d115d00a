"""The C++ front end (orz_b200/cli/orz_main.cpp) keeps src/main.rs's interface: flags, defaults, error strings."""
import os
import subprocess

import pytest

from oracle import pyoracle
from tests import datagen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORZ = os.path.join(ROOT, "orz_b200", "bin", "orz")


def _run(args, data=b""):
    return subprocess.run([ORZ] + args, input=data, capture_output=True, timeout=300)


def test_invalid_level_and_usage():
    if not os.path.exists(ORZ):
        pytest.skip("CLI not built")
    r = _run(["encode", "-l", "3"])
    assert r.returncode == 1 and b"invalid level: 3" in r.stderr  # main.rs:101
    assert _run([]).returncode == 2 and _run(["frobnicate"]).returncode == 2


def test_fails_loudly_without_gpu():
    import orz_b200
    if not os.path.exists(ORZ) or orz_b200.device_count() > 0:
        pytest.skip("needs the CLI and no GPU")
    r = _run(["encode", "-s"], b"hello")
    assert r.returncode == 1 and b"encoding failed" in r.stderr and r.stdout == b""  # main.rs:105
    r = _run(["decode", "-s"], b"\x00")
    assert r.returncode == 1 and b"decoding failed" in r.stderr  # main.rs:113


@pytest.mark.gpu
def test_cli_round_trip_matches_oracle(tmp_path):
    data = datagen.text_like(120_000, seed=77)
    src = tmp_path / "in.txt"
    src.write_bytes(data)
    for level in (0, 2):
        r = _run(["encode", "-s", "-l", str(level), str(src)])  # stdout
        assert r.returncode == 0 and r.stdout == pyoracle.encode(data, level)
    enc = tmp_path / "in.orz"
    r = _run(["encode", str(src), str(enc)])  # default level 2 (main.rs:30), statistics on stderr
    assert r.returncode == 0 and b"statistics:" in r.stderr and enc.read_bytes() == pyoracle.encode(data, 2)
    r = _run(["decode", "-s"], enc.read_bytes())  # stdin -> stdout
    assert r.returncode == 0 and r.stdout == data


@pytest.mark.gpu
def test_cli_segment_container(tmp_path):
    data = datagen.text_like(200_000, seed=78)
    src = tmp_path / "in.txt"
    src.write_bytes(data)
    r = _run(["encode", "-s", "-l", "1", "--segment-size", "65536", str(src)])
    assert r.returncode == 0 and r.stdout[:4] == b"ORZS"
    r2 = _run(["decode", "-s"], r.stdout)  # the container is recognised by its magic
    assert r2.returncode == 0 and r2.stdout == data
