"""Segment container (extension, include/orz_b200.h): layout on CPU, codec behaviour on the GPU."""
import zlib

import pytest

import orz_b200
from oracle import pyoracle
from tests import datagen


def _container_from_oracle(data: bytes, seg: int, level: int) -> bytes:
    segs = [data[i:i + seg] for i in range(0, len(data), seg)]
    streams = [pyoracle.encode(s, level) for s in segs]
    head = b"ORZS" + bytes([1, level, 0, 0]) + len(data).to_bytes(8, "little") + seg.to_bytes(4, "little") + len(segs).to_bytes(4, "little")
    table = b"".join(len(s).to_bytes(4, "little") + len(e).to_bytes(4, "little") + zlib.crc32(s).to_bytes(4, "little")
                     for s, e in zip(segs, streams))
    return head + table + b"".join(streams)


def test_container_layout_helpers():
    data = datagen.text_like(50_000, seed=3)
    c = _container_from_oracle(data, 16_384, 1)
    streams = orz_b200.segment_streams(c)
    assert len(streams) == 4
    assert b"".join(pyoracle.decode(s, 16_384 + 64) for s in streams) == data  # each segment is a plain orz stream
    assert orz_b200._lib.load().orz_segmented_bound(50_000, 16_384) >= len(c)


@pytest.mark.gpu
def test_segmented_matches_oracle_container_and_round_trips():
    data = datagen.text_like(300_000, seed=4) + datagen.random_bytes(5_000, seed=5)
    for seg, level in ((65_536, 2), (100_000, 0), (1 << 20, 1)):
        got = orz_b200.encode_segmented(data, seg, level)
        assert got == _container_from_oracle(data, seg, level)  # byte-identical to per-segment oracle streams
        assert orz_b200.decode_segmented(got) == data
    assert orz_b200.decode_segmented(orz_b200.encode_segmented(b"", 4096, 2)) == b""


@pytest.mark.gpu
def test_segmented_detects_corruption():
    data = datagen.text_like(120_000, seed=6)
    c = bytearray(orz_b200.encode_segmented(data, 32_768, 1))
    c[24 + 8] ^= 0xFF  # first segment's CRC
    with pytest.raises(orz_b200.OrzError) as e:
        orz_b200.decode_segmented(bytes(c))
    assert e.value.code == -1
    with pytest.raises(orz_b200.OrzError):
        orz_b200.decode_segmented(b"ORZX" + bytes(40))
