"""Host-side mirror of the reference's public interface for the ROLZ hot path.

Names and argument meaning follow richox/orz: `encode` / `decode` (src/lib.rs:58,94), `LZCfg`
(src/lz.rs:32-47), `LZEncoder` / `LZDecoder` with `new/encode/forward` and `new/decode/forward`
(src/lz.rs:75-95, 353-366).  All compute happens in the CUDA library behind include/orz_b200.h.
"""
import ctypes as C

import numpy as np

from ._lib import ORZ_ENOSPC, LZCfg, OrzError, check, load

LZ_BLOCK_SIZE = (1 << 25) - 1            # lib.rs:31
LZ_CHUNK_SIZE = 1 << 20                  # lib.rs:32
SBVEC_SENTINEL_LEN = 480                 # lib.rs:54
SBVEC_PREMATCH_LEN = LZ_BLOCK_SIZE // 2  # lib.rs:55
TBVEC_LEN = SBVEC_PREMATCH_LEN * 3       # lib.rs:68


def _u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8)
    return np.frombuffer(data, dtype=np.uint8)


def _cfg(cfg_or_level) -> LZCfg:
    return cfg_or_level if isinstance(cfg_or_level, LZCfg) else LZCfg.from_level(int(cfg_or_level))


def device_count() -> int:
    return load().orz_device_count()


def encode_bound(n: int) -> int:
    return load().orz_encode_bound(n)


def encode(source, cfg=2, device: int = 0) -> bytes:
    """orz::encode (lib.rs:58-92), memory to memory: returns the framed orz stream for `source`."""
    return encode_batch([source], cfg, device)[0]


def decode(target, out_cap: int = None, device: int = 0) -> bytes:
    """orz::decode (lib.rs:94-129), memory to memory."""
    return decode_batch([target], [out_cap] if out_cap is not None else None, device)[0]


def encode_batch(sources, cfg=2, device: int = 0):
    """Independent streams, one per thread block (extension; each equals `orz encode` of that input alone)."""
    L = load()
    cfg = _cfg(cfg)
    n = len(sources)
    ins = [_u8(s) for s in sources]
    caps = [L.orz_encode_bound(a.size) for a in ins]
    outs = [np.empty(c, dtype=np.uint8) for c in caps]
    inp = (C.c_void_p * n)(*[a.ctypes.data if a.size else None for a in ins])
    inl = (C.c_size_t * n)(*[a.size for a in ins])
    outp = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
    outc = (C.c_size_t * n)(*caps)
    outl = (C.c_size_t * n)()
    check(L.orz_encode_batch(device, C.byref(cfg), n, inp, inl, outp, outc, outl), "encoding failed")
    return [outs[i][: outl[i]].tobytes() for i in range(n)]


def encode_batch_multi(sources, cfg=2, devices=(0,)):
    """Independent streams sharded over several GPUs of this box behind the C ABI: stream i on devices[i % len(devices)]."""
    L = load()
    cfg = _cfg(cfg)
    n = len(sources)
    ins = [_u8(s) for s in sources]
    caps = [L.orz_encode_bound(a.size) for a in ins]
    outs = [np.empty(c, dtype=np.uint8) for c in caps]
    inp = (C.c_void_p * n)(*[a.ctypes.data if a.size else None for a in ins])
    inl = (C.c_size_t * n)(*[a.size for a in ins])
    outp = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
    outc = (C.c_size_t * n)(*caps)
    outl = (C.c_size_t * n)()
    devs = (C.c_int * len(devices))(*devices)
    check(L.orz_encode_batch_multi(devs, len(devices), C.byref(cfg), n, inp, inl, outp, outc, outl), "encoding failed")
    return [outs[i][: outl[i]].tobytes() for i in range(n)]


def decode_batch(targets, out_caps=None, device: int = 0):
    """Independent streams.  Without explicit capacities the output buffers start at 16x the compressed size and grow 4x on
    ORZ_ENOSPC (ROLZ streams of repetitive data exceed any fixed ratio; the reference writes to a growing Write)."""
    L = load()
    n = len(targets)
    ins = [_u8(t) for t in targets]
    grow = out_caps is None
    caps = list(out_caps) if out_caps is not None else [max(1 << 16, a.size * 16) for a in ins]
    for _ in range(8):
        outs = [np.empty(max(c, 1), dtype=np.uint8) for c in caps]
        inp = (C.c_void_p * n)(*[a.ctypes.data if a.size else None for a in ins])
        inl = (C.c_size_t * n)(*[a.size for a in ins])
        outp = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        outc = (C.c_size_t * n)(*caps)
        outl = (C.c_size_t * n)()
        rc = L.orz_decode_batch(device, n, inp, inl, outp, outc, outl)
        if rc == ORZ_ENOSPC and grow:
            caps = [c * 4 for c in caps]
            continue
        check(rc, "decoding failed")
        return [outs[i][: outl[i]].tobytes() for i in range(n)]
    check(ORZ_ENOSPC, "decoding failed")


def encode_segmented(source, segment_size: int = 1 << 24, cfg=2, device: int = 0) -> bytes:
    """Segment container (extension, include/orz_b200.h): independent orz streams of `segment_size` bytes + index + CRC32."""
    L = load()
    level = int(cfg) if not isinstance(cfg, LZCfg) else 255
    cfg = _cfg(cfg)
    a = _u8(source)
    cap = L.orz_segmented_bound(a.size, segment_size)
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t()
    check(L.orz_encode_segmented(device, C.byref(cfg), level, a.ctypes.data if a.size else None, a.size, segment_size,
                                 out.ctypes.data, cap, C.byref(n)), "encoding failed")
    return out[: n.value].tobytes()


def decode_segmented(target, device: int = 0) -> bytes:
    L = load()
    a = _u8(target)
    raw = C.c_size_t()
    check(L.orz_segmented_raw_len(a.ctypes.data, a.size, C.byref(raw)), "decoding failed")
    out = np.empty(max(raw.value, 1), dtype=np.uint8)
    n = C.c_size_t()
    check(L.orz_decode_segmented(device, a.ctypes.data, a.size, out.ctypes.data, raw.value, C.byref(n)), "decoding failed")
    return out[: n.value].tobytes()


def segment_streams(container) -> list:
    """The independent orz streams inside a container (each decodable by the reference `orz decode`)."""
    b = bytes(container)
    assert b[:4] == b"ORZS" and b[4] == 1
    nseg = int.from_bytes(b[20:24], "little")
    pos = 24 + 12 * nseg
    out = []
    for i in range(nseg):
        cl = int.from_bytes(b[24 + 12 * i + 4:24 + 12 * i + 8], "little")
        out.append(b[pos:pos + cl])
        pos += cl
    return out


class Batch:
    """Device-resident batch (include/orz_b200.h orz_batch_*): upload once, run() touches only HBM."""
    ENCODE, DECODE = 0, 1

    def __init__(self, mode, in_lens, out_caps=None, cfg=2, device: int = 0):
        L = load()
        self.n = len(in_lens)
        self.mode = mode
        self.cfg = _cfg(cfg)
        if out_caps is None:
            out_caps = [L.orz_encode_bound(n) for n in in_lens] if mode == 0 else [max(1 << 16, n * 16) for n in in_lens]
        self.out_caps = list(out_caps)
        il = (C.c_size_t * self.n)(*in_lens)
        oc = (C.c_size_t * self.n)(*self.out_caps)
        self.h = L.orz_batch_create(device, mode, C.byref(self.cfg), self.n, il, oc)
        if not self.h:
            raise OrzError(-5 if L.orz_device_count() == 0 else -4, "orz_batch_create")

    def close(self):
        if getattr(self, "h", None):
            load().orz_batch_free(self.h)
            self.h = None

    __del__ = close

    def upload(self, i, data):
        a = _u8(data)
        check(load().orz_batch_upload(self.h, i, a.ctypes.data if a.size else None, a.size), "upload")

    def upload_ptr(self, i, ptr, n):
        check(load().orz_batch_upload(self.h, i, ptr, n), "upload")

    def run(self):
        check(load().orz_batch_run(self.h), "run")

    def result_len(self, i) -> int:
        n = C.c_size_t()
        check(load().orz_batch_result_len(self.h, i, C.byref(n)), "result_len")
        return n.value

    def result_dev(self, i) -> int:
        return load().orz_batch_result_dev(self.h, i)

    def download(self, i) -> bytes:
        n = self.result_len(i)
        out = np.empty(max(n, 1), dtype=np.uint8)
        got = C.c_size_t()
        check(load().orz_batch_download(self.h, i, out.ctypes.data, n, C.byref(got)), "download")
        return out[: got.value].tobytes()

    def download_into(self, i, ptr, cap) -> int:
        got = C.c_size_t()
        check(load().orz_batch_download(self.h, i, ptr, cap, C.byref(got)), "download")
        return got.value

    def timing(self, cls):
        ms, k = C.c_float(), C.c_int()
        load().orz_batch_timing(self.h, cls, C.byref(ms), C.byref(k))
        return ms.value, k.value


class LZEncoder:
    """LZEncoder (lz.rs:69-95): codec state in HBM, chunks driven by the host like lib.rs:76-82."""

    def __init__(self, device: int = 0):
        self.h = load().orz_lzenc_new(device)
        if not self.h:
            raise OrzError(-5, "LZEncoder::new")

    def close(self):
        if getattr(self, "h", None):
            load().orz_lzenc_free(self.h)
            self.h = None

    __del__ = close

    def forward(self, forward_len: int):
        check(load().orz_lzenc_forward(self.h, forward_len), "forward")

    def encode(self, cfg, window: np.ndarray, sbuf_len: int, spos: int):
        """window: uint8[SENTINEL + LZ_BLOCK_SIZE + SENTINEL] (lib.rs:67); sbuf = window[SENTINEL:][:sbuf_len].
        Returns (new_spos, payload bytes) like lz.rs:89-95."""
        cfg = _cfg(cfg)
        assert window.dtype == np.uint8 and window.size >= SBVEC_SENTINEL_LEN * 2 + sbuf_len
        tbuf = np.empty(TBVEC_LEN, dtype=np.uint8)
        s, t = C.c_size_t(), C.c_size_t()
        check(load().orz_lzenc_encode(self.h, C.byref(cfg), window.ctypes.data + SBVEC_SENTINEL_LEN, sbuf_len,
                                      tbuf.ctypes.data, tbuf.size, spos, C.byref(s), C.byref(t)), "encode")
        return s.value, tbuf[: t.value].tobytes()


class LZDecoder:
    """LZDecoder (lz.rs:348-478)."""

    def __init__(self, device: int = 0):
        self.h = load().orz_lzdec_new(device)
        if not self.h:
            raise OrzError(-5, "LZDecoder::new")

    def close(self):
        if getattr(self, "h", None):
            load().orz_lzdec_free(self.h)
            self.h = None

    __del__ = close

    def forward(self, forward_len: int):
        check(load().orz_lzdec_forward(self.h, forward_len), "forward")

    def decode(self, tbuf, sbuf: np.ndarray, spos: int) -> int:
        """Decodes one chunk payload into sbuf[spos:spos_end]; sbuf is the LZ_BLOCK_SIZE-byte window (lib.rs:104)."""
        t = _u8(tbuf)
        end = C.c_size_t()
        check(load().orz_lzdec_decode(self.h, t.ctypes.data, t.size, sbuf.ctypes.data, sbuf.size, spos, C.byref(end)), "decode")
        return end.value


def encode_with_operators(source, cfg=2, device: int = 0) -> bytes:
    """lib.rs:58-92 written against LZEncoder exactly as the reference's host loop does (chunk-level ABI)."""
    cfg = _cfg(cfg)
    src = _u8(source)
    enc = LZEncoder(device)
    window = np.zeros(LZ_BLOCK_SIZE + 2 * SBVEC_SENTINEL_LEN, dtype=np.uint8)
    sbvec = window[SBVEC_SENTINEL_LEN:SBVEC_SENTINEL_LEN + LZ_BLOCK_SIZE]
    out = bytearray()
    ipos = 0
    while True:
        rd = min(src.size - ipos, LZ_BLOCK_SIZE - SBVEC_PREMATCH_LEN)
        if rd == 0:
            break
        sbvec[SBVEC_PREMATCH_LEN:SBVEC_PREMATCH_LEN + rd] = src[ipos:ipos + rd]
        ipos += rd
        spos = SBVEC_PREMATCH_LEN
        while spos < SBVEC_PREMATCH_LEN + rd:
            spos, payload = enc.encode(cfg, window, SBVEC_PREMATCH_LEN + rd, spos)
            out += _write_len(len(payload)) + payload
        sbvec[:SBVEC_PREMATCH_LEN] = sbvec[LZ_BLOCK_SIZE - SBVEC_PREMATCH_LEN:].copy()
        enc.forward(LZ_BLOCK_SIZE - SBVEC_PREMATCH_LEN)
    out += b"\x00"
    enc.close()
    return bytes(out)


def decode_with_operators(target, device: int = 0) -> bytes:
    """lib.rs:94-129 written against LZDecoder (chunk-level ABI)."""
    t = bytes(target)
    dec = LZDecoder(device)
    sbvec = np.zeros(LZ_BLOCK_SIZE, dtype=np.uint8)
    out = bytearray()
    pos, spos = 0, SBVEC_PREMATCH_LEN
    while True:
        n, pos = _read_len(t, pos)
        if n == 0:
            break
        if n >= TBVEC_LEN:
            raise OrzError(-1, "decoding failed")
        if pos + n > len(t):
            raise OrzError(-3, "decoding failed")
        end = dec.decode(t[pos:pos + n], sbvec, spos)
        pos += n
        out += sbvec[spos:end].tobytes()
        spos = end
        if spos >= LZ_BLOCK_SIZE:
            sbvec[:SBVEC_PREMATCH_LEN] = sbvec[LZ_BLOCK_SIZE - SBVEC_PREMATCH_LEN:].copy()
            dec.forward(LZ_BLOCK_SIZE - SBVEC_PREMATCH_LEN)
            spos = SBVEC_PREMATCH_LEN
    dec.close()
    return bytes(out)


def _write_len(n: int) -> bytes:  # ioutil.rs:79-88
    b = bytearray()
    while n >= 128:
        b.append(128 + n % 128)
        n //= 128
    b.append(n)
    return bytes(b)


def _read_len(buf: bytes, pos: int):  # ioutil.rs:60-76
    n, factor = 0, 1
    while True:
        if pos >= len(buf):
            raise OrzError(-3, "decoding failed")
        v = buf[pos]
        pos += 1
        if v < 128:
            return n + v * factor, pos
        n += (v - 128) * factor
        factor *= 128
