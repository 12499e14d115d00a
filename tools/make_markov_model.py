"""Builds data/markov2_enwik8.xz: exact order-2 byte counts of enwik8 restricted to {09,0A,20..7E}.

Run once in the authoring container (reads the reference's test corpus); the output is committed so that
bench.py and the tests can synthesise enwik8-shaped text on a box where /root/reference does not exist.
Format (little endian, lzma-compressed): u32 n | u16 delta[n] (gaps between non-zero trigram ids) | u32 count[n].
"""
import lzma
import sys

import numpy as np

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/test/enwik8.xz"
d = np.frombuffer(lzma.open(src).read(), dtype=np.uint8)
alpha = np.array([9, 10] + list(range(0x20, 0x7F)), dtype=np.uint8)
idx = np.full(256, -1, dtype=np.int16)
idx[alpha] = np.arange(97)
s = idx[d]
s = s[s >= 0].astype(np.int64)
tri = (s[:-2] * 97 + s[1:-1]) * 97 + s[2:]
counts = np.bincount(tri, minlength=97 ** 3)
nz = np.nonzero(counts)[0]
deltas = np.diff(nz, prepend=0)
assert deltas.max() < 65536
blob = np.uint32(len(nz)).tobytes() + deltas.astype("<u2").tobytes() + counts[nz].astype("<u4").tobytes()
open("data/markov2_enwik8.xz", "wb").write(lzma.compress(blob, preset=9))
print(len(nz), "trigrams ->", len(lzma.compress(blob, preset=9)), "bytes")
