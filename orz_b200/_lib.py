"""ctypes loader for the in-tree CUDA library (orz_b200/liborz_b200.so, built by __graft_entry__.build()).

There is no fallback: importing the codec without the built library, or running it without an sm_100
device, raises.  The oracle under oracle/ is never imported from here.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ORZ_B200_LIB") or os.path.join(_HERE, "liborz_b200.so")  # override: A/B builds of the same ABI

ORZ_OK, ORZ_EINVAL, ORZ_ENOSPC, ORZ_EEOF, ORZ_ENOMEM, ORZ_ECUDA, ORZ_EARG = 0, -1, -2, -3, -4, -5, -6


class LZCfg(C.Structure):
    """#[repr(C)] LZCfg (reference src/lz.rs:32-37)."""
    _fields_ = [("match_depth", C.c_size_t), ("lazy_match_depth1", C.c_size_t), ("lazy_match_depth2", C.c_size_t)]

    @classmethod
    def new(cls, match_depth, lazy_match_depth1, lazy_match_depth2):  # lz.rs:39-47
        return cls(match_depth, lazy_match_depth1, lazy_match_depth2)

    @classmethod
    def from_level(cls, level: int):  # main.rs:97-102
        cfg = cls()
        if load().orz_level_cfg(int(level), C.byref(cfg)) != ORZ_OK:
            raise ValueError(f"invalid level: {level}")
        return cfg


# every symbol include/orz_b200.h declares; tests check the library exports all of them
SYMBOLS = [
    "orz_level_cfg", "orz_strerror", "orz_device_count",
    "orz_lzenc_new", "orz_lzenc_free", "orz_lzenc_forward", "orz_lzenc_encode",
    "orz_lzdec_new", "orz_lzdec_free", "orz_lzdec_forward", "orz_lzdec_decode",
    "orz_encode_bound", "orz_encode_stream", "orz_decode_stream", "orz_encode_batch", "orz_decode_batch",
    "orz_encode_batch_multi", "orz_decode_batch_multi",
    "orz_batch_create", "orz_batch_free", "orz_batch_upload", "orz_batch_run", "orz_batch_result_len",
    "orz_batch_download", "orz_batch_result_dev", "orz_batch_timing",
    "orz_segmented_bound", "orz_encode_segmented", "orz_segmented_raw_len", "orz_decode_segmented",
]

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). orz_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    if hasattr(L, "orz_b200_is_emulator") and os.environ.get("ORZ_B200_TEST_EMULATOR") != "1":
        raise ImportError(f"{LIB_PATH} is the test-tier SIMT emulator build (tests/simt/), not the sm_100a library; "
                          "orz_b200 has no CPU fallback.")
    vp, sz, szp = C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)
    L.orz_level_cfg.argtypes = [C.c_int, C.POINTER(LZCfg)]
    L.orz_strerror.argtypes = [C.c_int]
    L.orz_strerror.restype = C.c_char_p
    L.orz_device_count.restype = C.c_int
    L.orz_lzenc_new.argtypes = [C.c_int]
    L.orz_lzenc_new.restype = vp
    L.orz_lzenc_free.argtypes = [vp]
    L.orz_lzenc_forward.argtypes = [vp, sz]
    L.orz_lzenc_encode.argtypes = [vp, C.POINTER(LZCfg), vp, sz, vp, sz, sz, szp, szp]
    L.orz_lzdec_new.argtypes = [C.c_int]
    L.orz_lzdec_new.restype = vp
    L.orz_lzdec_free.argtypes = [vp]
    L.orz_lzdec_forward.argtypes = [vp, sz]
    L.orz_lzdec_decode.argtypes = [vp, vp, sz, vp, sz, sz, szp]
    L.orz_encode_bound.argtypes = [sz]
    L.orz_encode_bound.restype = sz
    L.orz_encode_stream.argtypes = [C.c_int, C.POINTER(LZCfg), vp, sz, vp, sz, szp]
    L.orz_decode_stream.argtypes = [C.c_int, vp, sz, vp, sz, szp]
    L.orz_encode_batch.argtypes = [C.c_int, C.POINTER(LZCfg), C.c_int, C.POINTER(vp), szp, C.POINTER(vp), szp, szp]
    L.orz_decode_batch.argtypes = [C.c_int, C.c_int, C.POINTER(vp), szp, C.POINTER(vp), szp, szp]
    L.orz_encode_batch_multi.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(LZCfg), C.c_int, C.POINTER(vp), szp, C.POINTER(vp), szp, szp]
    L.orz_decode_batch_multi.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(vp), szp, C.POINTER(vp), szp, szp]
    L.orz_batch_create.argtypes = [C.c_int, C.c_int, C.POINTER(LZCfg), C.c_int, szp, szp]
    L.orz_batch_create.restype = vp
    L.orz_batch_free.argtypes = [vp]
    L.orz_batch_upload.argtypes = [vp, C.c_int, vp, sz]
    L.orz_batch_run.argtypes = [vp]
    L.orz_batch_result_len.argtypes = [vp, C.c_int, szp]
    L.orz_batch_download.argtypes = [vp, C.c_int, vp, sz, szp]
    L.orz_batch_result_dev.argtypes = [vp, C.c_int]
    L.orz_batch_result_dev.restype = vp
    L.orz_batch_timing.argtypes = [vp, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]
    L.orz_segmented_bound.argtypes = [sz, sz]
    L.orz_segmented_bound.restype = sz
    L.orz_encode_segmented.argtypes = [C.c_int, C.POINTER(LZCfg), C.c_int, vp, sz, sz, vp, sz, szp]
    L.orz_segmented_raw_len.argtypes = [vp, sz, szp]
    L.orz_decode_segmented.argtypes = [C.c_int, vp, sz, vp, sz, szp]
    _lib = L
    return L


class OrzError(IOError):
    """io::Error equivalent: .code is one of the ORZ_E* values of include/orz_b200.h."""

    def __init__(self, code, what=""):
        self.code = code
        msg = load().orz_strerror(code).decode()
        super().__init__(f"{what}: {msg}" if what else msg)


def check(code, what=""):
    if code != ORZ_OK:
        raise OrzError(code, what)
