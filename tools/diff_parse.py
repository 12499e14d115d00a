"""Where do two orz streams of the same input take different parse decisions?

    python tools/diff_parse.py GOT.bin WANT.bin [context_items]

Both streams are decoded chunk by chunk with the CPU oracle's chunk decoder (lib.rs:94-129 driven from here) with its item trace
on; the first chunk whose payload differs is compared item by item: position in the input, kind (literal / WORD / match), match
length and reduced offset.  Debugging tool for stage A (a timing-dependent protocol hole shows up as a different but valid parse).
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402


def read_len(t, pos):
    n, f = 0, 1
    while True:
        v = t[pos]
        pos += 1
        if v < 128:
            return n + v * f, pos
        n += (v - 128) * f
        f *= 128


class Dec:
    def __init__(self, stream):
        self.t = stream
        self.pos = 0
        self.L = po.lib()
        self.L.orz_oracle_dec_set_trace.argtypes = [C.c_void_p, C.c_size_t]
        self.L.orz_oracle_dec_trace_len.restype = C.c_size_t
        self.h = self.L.orz_oracle_dec_new()
        self.sb = np.zeros(po.BLOCK_SIZE + 1024, dtype=np.uint8)
        self.spos = po.PREMATCH
        self.base = 0  # input offset of window position PREMATCH
        self.trace = np.zeros(3 * (1 << 20), dtype=np.uint32)

    def next_chunk(self):
        """-> (payload bytes, trace rows [(input position, value, match length, reduced offset)]) or None at the end"""
        n, p = read_len(self.t, self.pos)
        if n == 0:
            return None
        payload = self.t[p:p + n]
        self.pos = p + n
        buf = np.frombuffer(payload + b"\0" * 16, dtype=np.uint8)
        self.L.orz_oracle_dec_set_trace(self.trace.ctypes.data, 1 << 20)
        end = C.c_size_t()
        rc = self.L.orz_oracle_dec_decode(self.h, buf.ctypes.data, n, self.sb.ctypes.data, self.spos, C.byref(end))
        k = self.L.orz_oracle_dec_trace_len()
        self.L.orz_oracle_dec_set_trace(None, 0)
        if rc != 0:
            raise RuntimeError(f"chunk at compressed offset {p} does not decode ({rc})")
        tr = self.trace[:3 * k].reshape(k, 3).copy()
        rows = np.stack([tr[:, 0].astype(np.int64) - po.PREMATCH + self.base, tr[:, 1] & 0xffff, tr[:, 1] >> 16, tr[:, 2]], axis=1)
        self.spos = end.value
        if self.spos >= po.BLOCK_SIZE:
            self.sb[:po.PREMATCH] = self.sb[po.BLOCK_SIZE - po.PREMATCH:po.BLOCK_SIZE].copy()
            self.L.orz_oracle_dec_forward(self.h, po.BLOCK_SIZE - po.PREMATCH)
            self.spos = po.PREMATCH
            self.base += po.BLOCK_SIZE - po.PREMATCH
        return payload, rows


def first_difference(got, want):
    """-> None, or (chunk index, item index in the chunk, items before the chunk, rows of got, rows of want)"""
    a, b = Dec(got), Dec(want)
    ci = before = 0
    while True:
        ca, cb = a.next_chunk(), b.next_chunk()
        if ca is None or cb is None:
            return None
        if ca[0] != cb[0]:
            ra, rb = ca[1], cb[1]
            m = min(len(ra), len(rb))
            neq = np.nonzero((ra[:m] != rb[:m]).any(axis=1))[0]
            return ci, (int(neq[0]) if len(neq) else m), before, ra, rb
        before += len(ca[1])
        ci += 1


def show(diff, ctx=6, out=sys.stdout):
    ci, k, before, ra, rb = diff
    print(f"chunk {ci}: {len(ra)} vs {len(rb)} items, first differing item {k} (item {before + k} of the stream)", file=out)
    for name, r in (("got", ra), ("want", rb)):
        print(f"  {name}:", file=out)
        for i in range(max(0, k - ctx), min(len(r), k + ctx)):
            pos, v, ml, ro = (int(x) for x in r[i])
            what = f"literal {v}" if v < 256 else (f"match len {ml} ro {ro} (sym {v})" if ml else f"WORD (sym {v})")
            print(f"   {'>' if i == k else ' '} item {i} pos {pos}: {what}", file=out)


def main():
    got = open(sys.argv[1], "rb").read()
    want = open(sys.argv[2], "rb").read()
    d = first_difference(got, want)
    if d is None:
        print("no chunk differs")
    else:
        show(d, int(sys.argv[3]) if len(sys.argv) > 3 else 6)


if __name__ == "__main__":
    main()
