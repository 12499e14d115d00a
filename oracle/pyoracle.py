"""ctypes binding of the CPU oracle (oracle/orz_oracle.c).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (orz_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liborz_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only); returns the .so path."""
    src = os.path.join(_HERE, "orz_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liborz_oracle.so", "orz_oracle"], stdout=subprocess.DEVNULL)
    return _SO


class Cfg(C.Structure):
    _fields_ = [("match_depth", C.c_size_t), ("lazy_match_depth1", C.c_size_t), ("lazy_match_depth2", C.c_size_t)]


class Item(C.Structure):
    _fields_ = [("symbol", C.c_uint16), ("ctx", C.c_uint16), ("robits", C.c_uint16), ("unlikely", C.c_uint8),
                ("robitlen", C.c_uint8), ("eml", C.c_uint8), ("after_literal", C.c_uint8), ("is_match", C.c_uint8),
                ("pad", C.c_uint8)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        u8p = C.POINTER(C.c_uint8)
        L.orz_oracle_level_cfg.argtypes = [C.c_int, C.POINTER(Cfg)]
        L.orz_oracle_encode_bound.argtypes = [C.c_size_t]
        L.orz_oracle_encode_bound.restype = C.c_size_t
        L.orz_oracle_encode_stream.argtypes = [C.POINTER(Cfg), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.orz_oracle_decode_stream.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.orz_oracle_enc_new.restype = C.c_void_p
        L.orz_oracle_enc_free.argtypes = [C.c_void_p]
        L.orz_oracle_enc_forward.argtypes = [C.c_void_p, C.c_size_t]
        L.orz_oracle_enc_encode.argtypes = [C.c_void_p, C.POINTER(Cfg), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                            C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.orz_oracle_enc_last_items.argtypes = [C.c_void_p, C.POINTER(C.POINTER(Item)), C.POINTER(C.POINTER(Item))]
        L.orz_oracle_enc_last_items.restype = C.c_size_t
        L.orz_oracle_dec_new.restype = C.c_void_p
        L.orz_oracle_dec_free.argtypes = [C.c_void_p]
        L.orz_oracle_dec_forward.argtypes = [C.c_void_p, C.c_size_t]
        L.orz_oracle_dec_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.orz_oracle_huffman_lens.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orz_oracle_huffman_codes.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orz_oracle_hash_dword.argtypes = [C.c_void_p]
        L.orz_oracle_hash_dword.restype = C.c_uint32
        L.orz_oracle_coder_selftest.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.orz_oracle_symrank_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        _lib = L
    return _lib


def level_cfg(level: int) -> Cfg:
    cfg = Cfg()
    if lib().orz_oracle_level_cfg(level, C.byref(cfg)) != 0:
        raise ValueError(f"invalid level: {level}")
    return cfg


def _as_u8(data) -> np.ndarray:
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    return np.ascontiguousarray(a, dtype=np.uint8)


def encode(data, level: int = 2, cfg: Cfg = None) -> bytes:
    """lib.rs:58-92 on the CPU oracle."""
    a = _as_u8(data)
    cfg = cfg or level_cfg(level)
    cap = lib().orz_oracle_encode_bound(a.size)
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t()
    rc = lib().orz_oracle_encode_stream(C.byref(cfg), a.ctypes.data, a.size, out.ctypes.data, cap, C.byref(n))
    if rc != 0:
        raise RuntimeError(f"oracle encode failed: {rc}")
    return out[: n.value].tobytes()


def decode(data, out_cap: int = None) -> bytes:
    """lib.rs:94-129 on the CPU oracle."""
    a = _as_u8(data)
    cap = out_cap if out_cap is not None else max(1 << 20, a.size * 16)
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t()
    rc = lib().orz_oracle_decode_stream(a.ctypes.data, a.size, out.ctypes.data, cap, C.byref(n))
    if rc != 0:
        raise RuntimeError(f"oracle decode failed: {rc}")
    return out[: n.value].tobytes()


def huffman_lens(weights, max_len: int = 15) -> np.ndarray:
    w = np.ascontiguousarray(weights, dtype=np.uint32)
    lens = np.zeros(w.size, dtype=np.uint8)
    lib().orz_oracle_huffman_lens(w.ctypes.data, w.size, max_len, lens.ctypes.data)
    return lens


def huffman_codes(lens) -> np.ndarray:
    l = np.ascontiguousarray(lens, dtype=np.uint8)
    codes = np.zeros(l.size, dtype=np.uint16)
    lib().orz_oracle_huffman_codes(l.ctypes.data, l.size, codes.ctypes.data)
    return codes


BLOCK_SIZE = (1 << 25) - 1
SENTINEL = 480
PREMATCH = BLOCK_SIZE // 2
TBUF_LEN = PREMATCH * 3


class ChunkEncoder:
    """LZEncoder (lz.rs:69-95) driven chunk by chunk, used to check the operator-level C ABI."""

    def __init__(self):
        self.h = lib().orz_oracle_enc_new()

    def __del__(self):
        if getattr(self, "h", None):
            lib().orz_oracle_enc_free(self.h)
            self.h = None

    def forward(self, n):
        lib().orz_oracle_enc_forward(self.h, n)

    def encode(self, cfg: Cfg, window: np.ndarray, sbuf_len: int, spos: int):
        """window: uint8 array of SENTINEL + BLOCK_SIZE + SENTINEL bytes; returns (spos, payload bytes)."""
        tbuf = np.empty(TBUF_LEN, dtype=np.uint8)
        s, t = C.c_size_t(), C.c_size_t()
        lib().orz_oracle_enc_encode(self.h, C.byref(cfg), window.ctypes.data + SENTINEL, sbuf_len, tbuf.ctypes.data, spos,
                                    C.byref(s), C.byref(t))
        return s.value, tbuf[: t.value].tobytes()

    def last_items(self):
        pre, post = C.POINTER(Item)(), C.POINTER(Item)()
        n = lib().orz_oracle_enc_last_items(self.h, C.byref(pre), C.byref(post))
        dt = np.dtype([("symbol", "<u2"), ("ctx", "<u2"), ("robits", "<u2"), ("unlikely", "u1"), ("robitlen", "u1"),
                       ("eml", "u1"), ("after_literal", "u1"), ("is_match", "u1"), ("pad", "u1")])
        a = np.ctypeslib.as_array(C.cast(pre, C.POINTER(C.c_uint8)), shape=(n * 12,)).view(dt).copy()
        b = np.ctypeslib.as_array(C.cast(post, C.POINTER(C.c_uint8)), shape=(n * 12,)).view(dt).copy()
        return a, b
