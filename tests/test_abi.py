"""The C-ABI library loads here (no GPU), exports every symbol include/orz_b200.h declares, and refuses to compute on CPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import orz_b200
from orz_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "orz_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(orz_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_lib.LIB_PATH)
    declared = _header_functions()
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/orz_b200.h but not exported"
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))


def test_level_cfg_and_bound():
    for level, want in [(0, (5, 3, 2)), (1, (15, 9, 6)), (2, (45, 27, 18))]:  # main.rs:97-102
        cfg = orz_b200.LZCfg.from_level(level)
        assert (cfg.match_depth, cfg.lazy_match_depth1, cfg.lazy_match_depth2) == want
    with pytest.raises(ValueError):
        orz_b200.LZCfg.from_level(3)
    assert orz_b200.encode_bound(0) >= 1 and orz_b200.encode_bound(10**8) > 10**8
    assert _lib.load().orz_strerror(-1) == b"invalid data"


def test_constants_match_reference():
    from orz_b200 import codec
    assert codec.LZ_BLOCK_SIZE == (1 << 25) - 1 and codec.SBVEC_PREMATCH_LEN == 16777215 and codec.TBVEC_LEN == 50331645
    assert codec._write_len(0) == b"\x00" and codec._write_len(300) == bytes([128 + 44, 2])  # ioutil.rs:79-88
    assert codec._read_len(bytes([128 + 44, 2]), 0) == (300, 2)


def test_no_cpu_fallback():
    if orz_b200.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(orz_b200.OrzError) as e:
        orz_b200.encode(b"hello world")
    assert e.value.code == -5  # ORZ_ECUDA
    with pytest.raises(orz_b200.OrzError):
        orz_b200.decode(b"\x00")
    with pytest.raises(orz_b200.OrzError):
        orz_b200.LZEncoder()


def test_product_package_never_imports_the_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "orz_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(root, f), errors="ignore").read()
                assert "pyoracle" not in text and "orz_oracle" not in text, f"{f} references the oracle"
