"""enwik8-shaped synthetic text: order-2 Markov chain with enwik8's own trigram counts (+1 smoothing)."""
import ctypes as C
import lzma
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
SEED = 0x6F727A5F62323030  # "orz_b200"
_model = None


def _load():
    global _model
    if _model is None:
        so = os.path.join(_HERE, "libtextgen.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
        lib = C.CDLL(so)
        lib.orz_textgen.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t]
        blob = lzma.open(os.path.join(ROOT, "data", "markov2_enwik8.xz")).read()
        n = int(np.frombuffer(blob[:4], "<u4")[0])
        deltas = np.frombuffer(blob[4:4 + 2 * n], "<u2").astype(np.int64)
        cnt = np.frombuffer(blob[4 + 2 * n:4 + 6 * n], "<u4").astype(np.int64)
        counts = np.ones(97 ** 3, dtype=np.int64)  # +1 smoothing
        counts[np.cumsum(deltas)] += cnt
        counts = counts.reshape(97 * 97, 97)
        cum = np.cumsum(counts, 1)
        thr = ((cum << 32) // cum[:, -1:] - 1).clip(0, 0xFFFFFFFF).astype(np.uint32)
        thr[:, -1] = 0xFFFFFFFF
        alpha = np.array([9, 10] + list(range(0x20, 0x7F)), dtype=np.uint8)
        _model = (lib, np.ascontiguousarray(thr), alpha)
    return _model


def markov_text(n: int, seed: int = SEED) -> np.ndarray:
    lib, thr, alpha = _load()
    out = np.empty(n, dtype=np.uint8)
    lib.orz_textgen(thr.ctypes.data, alpha.ctypes.data, seed, out.ctypes.data, n)
    return out


def corpus(n: int):
    """(data, description): real enwik8 (tiled if n > 100 MB) when data/enwik8.xz travelled with the repo, else Markov text."""
    p = os.path.join(ROOT, "data", "enwik8.xz")
    if os.path.exists(p):
        if n <= 100_000_000:  # a prefix: do not decompress the rest
            d = np.frombuffer(lzma.open(p).read(n), dtype=np.uint8)
            if d.size == n:
                return d.copy(), "enwik8 (data/enwik8.xz)"
        d = np.frombuffer(lzma.open(p).read(), dtype=np.uint8)
        if n <= d.size:
            return d[:n].copy(), "enwik8 (data/enwik8.xz)"
        return np.tile(d, (n + d.size - 1) // d.size)[:n].copy(), "enwik8 tiled (data/enwik8.xz)"
    return markov_text(n), "synthetic order-2 Markov text with enwik8 trigram counts (data/markov2_enwik8.xz)"
