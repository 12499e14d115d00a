/*
 * orz_b200.h — C ABI of the B200-native ROLZ codec (drop-in for richox/orz's hot path).
 *
 * The reference (richox/orz @29414c6, Rust) exposes no live FFI: `// pub mod ffi;` is commented out
 * at src/lib.rs:10.  This header defines the boundary a re-enabled `ffi.rs` would bind: one entry
 * point per reference operator, same argument meaning and error behaviour, plain pointers and sizes.
 * Every declaration cites the reference interface it replaces (file:line in the reference tree).
 * INTEGRATION.md shows the Rust-side `extern "C"` shim and a ctypes stub.
 *
 * All compute runs in hand-written sm_100a CUDA kernels.  There is no CPU fallback: every entry
 * point returns ORZ_ECUDA when no usable device / kernel image is present.
 */
#ifndef ORZ_B200_H
#define ORZ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes (io::ErrorKind equivalents of src/lib.rs:111-114, src/lz.rs:413-415) ---- */
#define ORZ_OK       0
#define ORZ_EINVAL  (-1) /* io::ErrorKind::InvalidData  (lib.rs:111-113, lz.rs:413-415) */
#define ORZ_ENOSPC  (-2) /* caller's output buffer too small (Write would fail)           */
#define ORZ_EEOF    (-3) /* io::ErrorKind::UnexpectedEof (lib.rs:114 read_exact)          */
#define ORZ_ENOMEM  (-4) /* host or device allocation failed                              */
#define ORZ_ECUDA   (-5) /* no device / kernel launch failure — never falls back to CPU   */
#define ORZ_EARG    (-6) /* bad argument (e.g. level outside 0..2: main.rs:101)           */

/* ---- constants that are part of the contract (src/lib.rs:31-34,54-55) ---- */
#define ORZ_BLOCK_SIZE    ((1u << 25) - 1) /* LZ_BLOCK_SIZE                                   */
#define ORZ_CHUNK_ITEMS   (1u << 20)       /* LZ_CHUNK_SIZE: items per chunk                  */
#define ORZ_SENTINEL_LEN  480u             /* SBVEC_SENTINEL_LEN: readable bytes either side  */
#define ORZ_PREMATCH_LEN  (ORZ_BLOCK_SIZE / 2) /* SBVEC_PREMATCH_LEN: first stream position   */
#define ORZ_TBUF_LEN      (ORZ_PREMATCH_LEN * 3) /* lib.rs:68,103 payload buffer length       */

/* src/lz.rs:32-37 — #[repr(C)] pub struct LZCfg */
typedef struct {
    size_t match_depth;
    size_t lazy_match_depth1;
    size_t lazy_match_depth2;
} orz_lzcfg;

/* src/main.rs:97-102 — level 0/1/2 -> LZCfg (5,3,2) / (15,9,6) / (45,27,18); ORZ_EARG otherwise */
int orz_level_cfg(int level, orz_lzcfg *cfg);

const char *orz_strerror(int code);
/* number of usable sm_100 devices (0 if none); never throws */
int orz_device_count(void);

/* =====================================================================================
 * Operator interface — replaces LZEncoder / LZDecoder (callers: src/lib.rs:64,78,84,99,116,122)
 * ===================================================================================== */
typedef struct orz_lzenc orz_lzenc;
typedef struct orz_lzdec orz_lzdec;

/* src/lz.rs:75-80  LZEncoder::new() — all codec state lives in HBM on `device` */
orz_lzenc *orz_lzenc_new(int device);
void orz_lzenc_free(orz_lzenc *);
/* src/lz.rs:82-87  LZEncoder::forward(forward_len) — rebases ring positions, unlinks dead chain entries */
int orz_lzenc_forward(orz_lzenc *, size_t forward_len);
/* src/lz.rs:89-95  LZEncoder::encode(cfg, sbuf, tbuf, spos) -> (spos, payload_len)
 * sbuf[-480 .. sbuf_len+480) must be readable (lib.rs:54,67-69): those bytes are bit-visible.
 * The window is (re)uploaded when spos == ORZ_PREMATCH_LEN or the (sbuf, sbuf_len, spos) sequence
 * is not a continuation of the previous call; lib.rs:76-82 never modifies sbuf between the chunks
 * of one block. */
int orz_lzenc_encode(orz_lzenc *, const orz_lzcfg *cfg, const uint8_t *sbuf, size_t sbuf_len, uint8_t *tbuf,
                     size_t tbuf_cap, size_t spos, size_t *spos_out, size_t *tlen_out);

/* src/lz.rs:353-357  LZDecoder::new() */
orz_lzdec *orz_lzdec_new(int device);
void orz_lzdec_free(orz_lzdec *);
/* src/lz.rs:359-364  LZDecoder::forward(forward_len) */
int orz_lzdec_forward(orz_lzdec *, size_t forward_len);
/* src/lz.rs:366  LZDecoder::decode(tbuf, sbuf, spos) -> io::Result<spos_end>
 * Decodes one chunk payload into sbuf[spos..spos_end) (host memory; sbuf is the lib.rs:102-104
 * window of ORZ_BLOCK_SIZE bytes).  Returns ORZ_OK or ORZ_EINVAL.  Unlike the reference (unchecked
 * indexing, lib.rs:36-39) malformed input is bounds-checked instead of being undefined behaviour. */
int orz_lzdec_decode(orz_lzdec *, const uint8_t *tbuf, size_t tlen, uint8_t *sbuf, size_t sbuf_cap, size_t spos,
                     size_t *spos_end);

/* =====================================================================================
 * Stream interface — replaces orz::encode / orz::decode (src/lib.rs:58-92, 94-129), memory to memory.
 * Output of orz_encode_stream is byte-identical to the in-tree restatements of `orz encode -l<level>` (oracle/orz_oracle.c and the
 * independent oracle/pyref.py agree on it); parity with bytes produced by a real orz build is NOT pinned yet: no Rust toolchain exists
 * where this was written (tools/make_ref_golden.sh produces the pin wherever nightly Rust is available).
 * ===================================================================================== */
size_t orz_encode_bound(size_t n);
int orz_encode_stream(int device, const orz_lzcfg *cfg, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap,
                      size_t *out_len);
int orz_decode_stream(int device, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap, size_t *out_len);

/* =====================================================================================
 * Batch interface (extension; the reference is single-stream, single-thread).
 * Each stream is an independent orz stream (fresh LZEncoder, lib.rs:64) and is byte-identical to the
 * reference run on that input alone.  Streams run concurrently on one GPU, one per thread block.
 * ===================================================================================== */
int orz_encode_batch(int device, const orz_lzcfg *cfg, int nstreams, const uint8_t *const *in, const size_t *n,
                     uint8_t *const *out, const size_t *out_cap, size_t *out_len);
int orz_decode_batch(int device, int nstreams, const uint8_t *const *in, const size_t *n, uint8_t *const *out,
                     const size_t *out_cap, size_t *out_len);

/* Several GPUs of one box behind the C ABI (no torch, no NCCL needed by the caller): stream i runs on devices[i % ndev], one host
 * thread per device, results land in the caller's `out[i]` (the "gather" is the D2H copy of each span).  Streams are independent orz
 * streams exactly as in orz_encode_batch; one orz stream does not shard.  devices[] may repeat an index. */
int orz_encode_batch_multi(const int *devices, int ndev, const orz_lzcfg *cfg, int nstreams, const uint8_t *const *in, const size_t *n,
                           uint8_t *const *out, const size_t *out_cap, size_t *out_len);
int orz_decode_batch_multi(const int *devices, int ndev, int nstreams, const uint8_t *const *in, const size_t *n, uint8_t *const *out,
                           const size_t *out_cap, size_t *out_len);

/* Device-resident batch: inputs are uploaded once, run() touches only HBM (what bench.py's `value` times),
 * download() brings results back.  mode: 0 = encode, 1 = decode. */
typedef struct orz_batch orz_batch;
orz_batch *orz_batch_create(int device, int mode, const orz_lzcfg *cfg, int nstreams, const size_t *in_len,
                            const size_t *out_cap);
void orz_batch_free(orz_batch *);
int orz_batch_upload(orz_batch *, int stream, const uint8_t *host, size_t n);
int orz_batch_run(orz_batch *);
int orz_batch_result_len(orz_batch *, int stream, size_t *len);
int orz_batch_download(orz_batch *, int stream, uint8_t *host, size_t cap, size_t *len);
/* device pointer of stream i's result (for NCCL gathers of compressed spans) */
void *orz_batch_result_dev(orz_batch *, int stream);

/* =====================================================================================
 * Segment container (extension; the reference format has no index, size or checksum).
 *   "ORZS" u8 version=1 u8 level u16 0 | u64 raw_len | u32 segment_size | u32 n_segments |
 *   n_segments x { u32 raw_len, u32 comp_len, u32 crc32(raw) } | the n_segments orz streams back to back
 * Every segment is an independent orz stream, byte-identical to `orz encode` of that segment alone, so segments encode
 * and decode concurrently (one thread block each) and shard across GPUs.  A reference `orz decode` fed one stream
 * extracted from the container decodes that segment.  All integers little endian.
 * ===================================================================================== */
size_t orz_segmented_bound(size_t n, size_t segment_size);
int orz_encode_segmented(int device, const orz_lzcfg *cfg, int level_tag, const uint8_t *in, size_t n, size_t segment_size,
                         uint8_t *out, size_t out_cap, size_t *out_len);
/* raw size recorded in a container header (ORZ_EINVAL if `in` is not a container) */
int orz_segmented_raw_len(const uint8_t *in, size_t n, size_t *raw_len);
int orz_decode_segmented(int device, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap, size_t *out_len);

/* Kernel timing of the last run(): CUDA-event milliseconds and launch counts per kernel class
 * (0 parse/stage A, 1 emit/stages B-D, 2 block load/slide/forward, 3 decode, 4 = the whole run() on the
 * library's stream, state reset included).  Returns launches counted. */
int orz_batch_timing(orz_batch *, int kernel_class, float *total_ms, int *launches);

#ifdef __cplusplus
}
#endif
#endif
