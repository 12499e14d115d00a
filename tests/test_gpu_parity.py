"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit-exact."""
import numpy as np
import pytest

import orz_b200
from oracle import pyoracle
from tests import datagen

pytestmark = pytest.mark.gpu


def _first_diff(a: bytes, b: bytes) -> str:
    n = min(len(a), len(b))
    x = np.frombuffer(a[:n], np.uint8) != np.frombuffer(b[:n], np.uint8)
    i = int(np.argmax(x)) if x.any() else n
    return f"len {len(a)} vs {len(b)}, first diff at byte {i}: {a[i:i+8].hex()} vs {b[i:i+8].hex()}"


@pytest.mark.parametrize("data", [b"", b"a", b"ab", b"abc", b"abcd", b"abcde", b"aaaaaaaaaaaaaaaaaaaaaaaaaaaa",
                                  b"the quick brown fox jumps over the lazy dog. " * 40])
@pytest.mark.parametrize("level", [0, 1, 2])
def test_tiny_inputs(data, level):
    want = pyoracle.encode(data, level)
    got = orz_b200.encode(data, level)
    assert got == want, _first_diff(got, want)
    assert orz_b200.decode(want, out_cap=len(data) + 64) == data


@pytest.mark.parametrize("level", [0, 1, 2])
@pytest.mark.parametrize("gen,n", [("text_like", 300_000), ("random_bytes", 70_000), ("low_entropy", 400_000)])
def test_encode_matches_oracle(gen, n, level):
    data = getattr(datagen, gen)(n)
    want = pyoracle.encode(data, level)
    got = orz_b200.encode(data, level)
    assert got == want, _first_diff(got, want)


@pytest.mark.parametrize("gen,n", [("text_like", 300_000), ("random_bytes", 70_000), ("low_entropy", 400_000)])
def test_decode_matches_oracle(gen, n):
    data = getattr(datagen, gen)(n)
    comp = pyoracle.encode(data, 1)
    assert orz_b200.decode(comp, out_cap=n + 64) == data
