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
@pytest.mark.parametrize("gen,n", [("text_like", 300_000), ("random_bytes", 70_000), ("low_entropy", 400_000), ("long_chains", 400_000)])
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


def test_gpu_matches_committed_golden_vectors():
    """The committed oracle vectors (tests/golden/golden.json): the GPU must reproduce every hash."""
    import hashlib
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))
    cases = {"empty": b"", "a": b"a", "abcabc": b"abcabcabcabcabcabcabcabc", "fox": b"the quick brown fox jumps over the lazy dog. " * 20,
             "text_64k": datagen.text_like(65536, seed=101), "random_20k": datagen.random_bytes(20000, seed=102),
             "lowent_100k": datagen.low_entropy(100000, seed=103)}
    for level in (0, 1, 2):
        outs = orz_b200.encode_batch(list(cases.values()), level)
        for (name, _), enc in zip(cases.items(), outs):
            want = g["cases"][name]["levels"][str(level)]
            assert (len(enc), hashlib.sha256(enc).hexdigest()) == (want["len"], want["sha256"]), (name, level)
    e8 = datagen.enwik8(2_000_000)
    if e8 is not None:
        enc = orz_b200.encode(e8, 2)
        want = g["enwik8_2M"]["levels"]["2"]
        assert (len(enc), hashlib.sha256(enc).hexdigest()) == (want["len"], want["sha256"])
