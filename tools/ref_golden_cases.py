"""The inputs of the reference golden vectors (tools/make_ref_golden.sh writes them to files and runs the real `orz encode`;
tests/test_oracle.py regenerates them to check the oracle against the recorded hashes).  `python tools/ref_golden_cases.py DIR`
writes every case to DIR/<name>.bin."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import datagen  # noqa: E402

CASES = ["empty", "a", "abcd", "fox", "match_vector", "lazy_vector", "text_300k", "lowent_300k", "random_50k", "long_chains_200k",
         "enwik8_2M", "enwik8_20M", "two_blocks_34M"]


def case_bytes(name: str) -> bytes:
    if name == "empty":
        return b""
    if name == "a":
        return b"a"
    if name == "abcd":
        return b"abcd"
    if name == "fox":
        return b"the quick brown fox jumps over the lazy dog. " * 30
    if name == "match_vector":
        return b"0123456789abcdef" * 2 + b"XYZW"
    if name == "lazy_vector":
        return b".9ABCD!~-zABCDEFGH?~,9ABCDEFGH#qwertyuiop"
    if name == "text_300k":
        return datagen.text_like(300_000, seed=201)
    if name == "lowent_300k":
        return datagen.low_entropy(300_000, seed=202)
    if name == "random_50k":
        return datagen.random_bytes(50_000, seed=203)
    if name == "long_chains_200k":
        return datagen.long_chains(200_000, seed=204)
    if name == "enwik8_2M":
        return datagen.enwik8(2_000_000)
    if name == "enwik8_20M":  # one block slide + forward (lib.rs:83-84)
        return datagen.enwik8(20_000_000)
    if name == "two_blocks_34M":  # a short final block: stale tail bytes of the previous block are bit-visible
        return (datagen.text_like(3_000_000, seed=205) + datagen.low_entropy(1_250_000, seed=206)) * 8
    raise KeyError(name)


if __name__ == "__main__":
    out = sys.argv[1]
    os.makedirs(out, exist_ok=True)
    for c in CASES:
        d = case_bytes(c)
        if d is None:
            print("skip", c, "(data/enwik8.xz missing)")
            continue
        open(os.path.join(out, c + ".bin"), "wb").write(d)
