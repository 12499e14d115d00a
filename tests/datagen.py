"""Seeded synthetic inputs for the parity tests (no file under /root/reference is read at test time)."""
import lzma
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def text_like(n: int, seed: int = 1, vocab: int = 4000) -> bytes:
    """Zipf-distributed words over a random vocabulary with punctuation and newlines: LZ-friendly, text-shaped."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(1, 11, size=vocab)
    words = [bytes(rng.integers(97, 123, size=int(l), dtype=np.uint8)) for l in lens]
    seps = [b" ", b" ", b" ", b" ", b", ", b". ", b"\n", b" [[", b"]] ", b" &amp; "]
    out = bytearray()
    while len(out) < n:
        ids = rng.zipf(1.25, size=20000) % vocab
        sp = rng.integers(0, len(seps), size=20000)
        for i, s in zip(ids, sp):
            out += words[int(i)]
            out += seps[int(s)]
    return bytes(out[:n])


def random_bytes(n: int, seed: int = 2) -> bytes:
    return np.random.default_rng(seed).integers(0, 256, size=n, dtype=np.uint8).tobytes()


def low_entropy(n: int, seed: int = 3) -> bytes:
    """Long runs and short periods: exercises 240-byte matches, overlapped copies and ring wrap-around."""
    rng = np.random.default_rng(seed)
    out = bytearray()
    while len(out) < n:
        kind = int(rng.integers(0, 3))
        if kind == 0:
            out += bytes([int(rng.integers(0, 256))]) * int(rng.integers(1, 2000))
        elif kind == 1:
            p = bytes(rng.integers(0, 256, size=int(rng.integers(1, 9)), dtype=np.uint8))
            out += p * int(rng.integers(1, 300))
        else:
            out += bytes(rng.integers(32, 127, size=int(rng.integers(1, 50)), dtype=np.uint8))
    return bytes(out[:n])


def long_chains(n: int, seed: int = 4) -> bytes:
    """Three 5-byte strings repeated thousands of times with different continuations: every occurrence lands on the same
    (context, hash entry) chain, so chain walks run to the depth limit, past the 29 links of a ring-record snapshot
    (multi-round gathers, snapshot source switches) while the matches themselves stay short."""
    rng = np.random.default_rng(seed)
    pats = [b" zqxj", b" kwvb", b"-mmmm"]
    out = bytearray()
    while len(out) < n:
        out += pats[int(rng.integers(0, 3))] + bytes(rng.integers(97, 123, size=int(rng.integers(1, 4)), dtype=np.uint8))
        if rng.random() < 0.3:
            out += bytes(rng.integers(32, 127, size=int(rng.integers(1, 12)), dtype=np.uint8))
    return bytes(out[:n])


def enwik8(n: int = None, offset: int = 0):
    """First n bytes of enwik8 if data/enwik8.xz travelled with the repo, else None."""
    p = os.path.join(ROOT, "data", "enwik8.xz")
    if not os.path.exists(p):
        return None
    with lzma.open(p) as f:
        if offset:
            f.seek(offset)
        return f.read(n) if n is not None else f.read()
