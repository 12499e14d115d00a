// orz_dev.cuh — shared device-side definitions for the B200 ROLZ codec kernels.
//
// Data layout in HBM (per stream) and the constants of the wire format.  Reference citations are
// file:line into richox/orz @29414c6.  Nothing here is a translation of the Rust structs: the ring
// node, chain link and "expected length" byte that the reference keeps in three arrays
// (matcher.rs:28-31,102-105) are fused into one 8-byte record so that one chain hop is one load.
#pragma once
#include <cstdint>
#ifdef ORZ_SIMT_EMU
// Test tier only: tests/simt/ compiles these same sources for the host-side SIMT emulator (fibres instead of
// CUDA threads) so that the CPU-only tests can execute the kernel source against the oracle.  The product
// library is always built by nvcc without this macro.
#include "simt_rt.h"
#define ORZ_LAUNCH(kern, grid, block, smem, st, ...) simt::launch(dim3(grid), dim3(block), (smem), [=]() { kern(__VA_ARGS__); })
#define ORZ_LAUNCH_CLUSTER(kern, grid, block, smem, st, k, ...) \
    simt::launch(dim3(grid), dim3(block), (smem), [=]() { kern(__VA_ARGS__); }, (k))
#define ORZ_COST(c) simt::cost(c)   // virtual-clock annotation at a dependent load (latency model, SIMT_VT=1)
#define ORZ_POLL() simt::poll()     // a spin-loop iteration: let the other fibres run
#else
#include <cuda_runtime.h>
#define ORZ_DYN_SHARED(name) extern __shared__ __align__(16) uint8_t name[]
#define ORZ_LAUNCH(kern, grid, block, smem, st, ...) kern<<<(grid), (block), (smem), (st)>>>(__VA_ARGS__)
// thread-block cluster of k CTAs along x (grid must be a multiple of k)
#define ORZ_LAUNCH_CLUSTER(kern, grid, block, smem, st, k, ...)                                                     \
    do {                                                                                                            \
        cudaLaunchConfig_t cfg_ = {};                                                                               \
        cfg_.gridDim = dim3(grid);                                                                                  \
        cfg_.blockDim = dim3(block);                                                                                \
        cfg_.dynamicSmemBytes = (smem);                                                                             \
        cfg_.stream = (st);                                                                                         \
        cudaLaunchAttribute at_[1];                                                                                 \
        at_[0].id = cudaLaunchAttributeClusterDimension;                                                            \
        at_[0].val.clusterDim.x = (k);                                                                              \
        at_[0].val.clusterDim.y = 1;                                                                                \
        at_[0].val.clusterDim.z = 1;                                                                                \
        cfg_.attrs = at_;                                                                                           \
        cfg_.numAttrs = 1;                                                                                          \
        cudaLaunchKernelEx(&cfg_, kern, __VA_ARGS__);                                                               \
    } while (0)
#define ORZ_COST(c) ((void)0)
#define ORZ_POLL() ((void)0)
#endif

namespace orz {

// ---- wire-format / algorithm constants -------------------------------------------------------
#if defined(ORZ_SIMT_EMU) && defined(ORZ_TEST_BLOCK_BITS)
// Emulator-only scale model (tests/simt/): small blocks and chunks so that the CPU tier reaches block slides,
// `forward` and chunk cuts within kilobytes.  The oracle is built with the same two values for those tests.
constexpr uint32_t kBlockSize = (1u << ORZ_TEST_BLOCK_BITS) - 1;
constexpr uint32_t kChunkItems = ORZ_TEST_CHUNK_ITEMS;
#else
constexpr uint32_t kBlockSize = (1u << 25) - 1;        // lib.rs:31  LZ_BLOCK_SIZE
constexpr uint32_t kChunkItems = 1u << 20;             // lib.rs:32  LZ_CHUNK_SIZE
#endif
constexpr int kMatchMax = 240;                         // lib.rs:33
constexpr int kMatchMin = 4;                           // lib.rs:34
constexpr uint32_t kSentinel = 480;                    // lib.rs:54
constexpr uint32_t kPrematch = kBlockSize / 2;         // lib.rs:55  (16,777,215)
constexpr uint32_t kBlockNew = kBlockSize - kPrematch; // new bytes per block (16,777,216)
constexpr uint32_t kTbufLen = kPrematch * 3;           // lib.rs:68
constexpr int kRing = 4094;                            // lz.rs:24   LZ_MF_BUCKET_ITEM_SIZE
constexpr int kHashSize = 4627;                        // matcher.rs:18
constexpr int kRoidSize = 22;                          // lz.rs:27
constexpr int kLenidSize = 6;                          // lz.rs:28
constexpr int kNumSym = 256 + kRoidSize * kLenidSize + 1; // lz.rs:25 (389)
constexpr int kWordSym = kNumSym - 1;                  // lz.rs:29
constexpr int kNumBuckets = 256;
constexpr int kNumCtx = 512;

// ---- HBM layout ---------------------------------------------------------------------------------
constexpr int kNodesStride = 4096;  // ring slots per bucket, padded from 4094
constexpr int kHeadsStride = 4640;  // hash heads per bucket, padded from 4627
constexpr int kSymStride = 392;     // symbols per rank table, padded from 389
constexpr uint32_t kWinSlack = 8192; // zero bytes kept readable after the window (tile loads)

// Decoder ring slot: pos:25 | match_len_min:7 (matcher.rs:231-237), expected (matcher.rs:239-243); `next` unused
struct __align__(8) Node {
    uint32_t pl;
    int16_t next;
    uint8_t expected;
    uint8_t pad;
};
static_assert(sizeof(Node) == 8, "Node must be 8 bytes");

// Encoder ring slot, 64 bytes.  Besides pos|lenmin and expected it carries `next[0]` = the reference's chain
// link nexts[slot] (matcher.rs:104,119) and `next[1..]` = a snapshot, taken when the slot was written, of the
// following links of that chain.  A walk (matcher.rs:149-181) then needs one gather of up to kSnap records
// instead of one dependent load per hop.  The snapshot may go stale, exactly like the reference's own links:
// a slot that was overwritten after the snapshot holds a newer position, and the walk's "positions strictly
// decrease" test (matcher.rs:176-179) stops there in both formulations.
constexpr int kSnap = 29;
struct __align__(64) ENode {
    uint32_t pl;
    uint8_t expected;
    uint8_t pad;
    int16_t next[kSnap];
};
static_assert(sizeof(ENode) == 64, "ENode must be 64 bytes");

__host__ __device__ __forceinline__ uint32_t node_pos(uint32_t pl) { return pl & 0x1ffffffu; }
__host__ __device__ __forceinline__ uint32_t node_lenmin(uint32_t pl) { return pl >> 25; }

// One parsed step (lz.rs:100-117 MatchItem), 16 bytes so a step is one st.v4
struct __align__(16) Item {
    uint16_t symbol;   // pre-rank symbol after stage A, rank after stage B
    uint16_t ctx;      // symrank_context (9 bits)
    uint16_t robits;
    uint8_t unlikely;
    uint8_t robitlen;
    uint8_t eml;       // encoded_match_len
    uint8_t flags;     // bit0 after_literal, bit1 is_match
    uint16_t pad;
    uint32_t seq;      // ordinal of this item among the chunk's items of the same ctx
};
static_assert(sizeof(Item) == 16, "Item must be 16 bytes");

// symrank.rs:14-19
struct SymRank {
    uint16_t value[kSymStride];
    uint16_t index[kSymStride];
    uint32_t cnt;
    uint32_t sum;
};
static_assert(sizeof(SymRank) == 1576, "SymRank size");

// Per-stream encoder control block (lz.rs:49-80 LZContext+LZEncoder, lib.rs:64-90 loop state)
struct EncStream {
    // set by host
    const uint8_t *in;    // whole input, HBM resident
    uint64_t in_len;
    uint8_t *out;         // framed orz stream
    uint64_t out_cap;
    uint8_t *win;         // win[p] is window position p (p may start at win_lo > 0 for one-block streams); 16 B aligned at p=0
    uint32_t win_lo, win_hi;
    ENode *nodes;         // [256][kNodesStride]
    int16_t *heads;       // [256][kHeadsStride]
    uint16_t *words;      // [32768] two predicted bytes, little endian (lz.rs:52)
    SymRank *symranks;    // [512]
    Item *items;          // [items_cap]
    uint32_t *sorted;     // [items_cap] item indices grouped by ctx
    uint8_t *payload;     // chunk payload scratch, 4 B aligned
    uint32_t payload_cap;
    uint32_t items_cap;
    uint32_t depth0, depth1, depth2; // LZCfg (lz.rs:32-37)
    int32_t self_driven;  // 1: stream mode (kernel runs the lib.rs loop), 0: chunk mode (host drives)
    // stage A v7 (parse_v7.cuh): per-key candidate lists instead of ring records + hash chains
    uint2 *klist;         // [256 * kHeadsStride][64] circular candidate lists: pos | expected, bucket ordinal
    uint32_t *kcnt;       // [256 * kHeadsStride] entries ever appended to the list
    uint8_t *lm;          // [256][kNodesStride] match_len_min of the node in each ring slot (matcher.rs:231-237)
    uint32_t *gext;       // [512][48] candidate positions of the probe records in flight
    uint32_t v7;          // 1: this stream uses the v7 layout (nodes / heads are not allocated)
    uint32_t need_slide;  // set at the end of a block: the advance kernels slide, forward and load
    Item *items_alt;      // v7: second item buffer — stages B-D of chunk k read one while stage A of chunk k+1 fills the other
    uint32_t dbg_flags;   // measurement switches (ORZ_B200_DBG): 1 no re-stamp, 8 executor cycle counters, 32 builder counters, 64 v3 idle builders sleep
    // mutable
    uint32_t bhead[kNumBuckets];
    uint32_t bcnt[kNumBuckets];  // v7: inserts into each bucket so far (ring head = bcnt % 4094)
    uint32_t ctx_count[kNumCtx];
    uint32_t spos, sbuf_len, after_literal, first_block, n_items;
    uint32_t payload_len;
    uint64_t in_pos, out_pos;
    int32_t status;       // 0 running, 1 done, <0 error code
    uint32_t parse_done;  // v7: stage A reached the end of the input (the last chunk may still be in stages B-D)
    uint32_t n_chunks;
    uint32_t n_items_alt;
    uint32_t chunk_end_spos[2]; // v7, per item buffer: min(spos, sbuf_len) after the chunk (lz.rs:268)
    uint32_t chunk_flags[2];    // v7, per item buffer: bit 0 = the stream ends after this chunk (lib.rs:89)
    uint32_t ctx_count_alt[kNumCtx];
    unsigned long long dbg[8]; // stage A counters: steps, executor spin iterations (own / lazy1 / lazy2 probe), builds, aborted builds
    unsigned long long dbg7[24]; // v7: resolver cycles total, load, walk, hand-over, repair, wait; records repaired, rebuilt from HBM; builder cycles; list lag
    // event trace of one position (instrumented build only; ORZ_B200_DBG_POS = window position, ORZ_B200_DBG_INPOS = input bytes
    // consumed when its block is loaded): builds, repairs, decisions and the hand-over of the records at that position and the two after it
    unsigned long long dbg_inpos;
    uint32_t dbg_pos;
    uint32_t dbgx_n;
    uint32_t dbgx[1024];
    // provenance of the probe records (A/B builds with ORZ_V7_PROV, ORZ_B200_PROV=1): four uint4 per input position (build, build, last repair, what the repair read)
    uint4 *prov;
    uint32_t prov_cap, prov_n;
};

// Per-stream decoder control block (lz.rs:348-357 LZDecoder, lib.rs:99-127 loop state)
struct DecStream {
    const uint8_t *in;    // framed orz stream (stream mode) or one chunk payload (chunk mode)
    uint64_t in_len;
    uint8_t *out;         // stream mode: out[0] is the first decoded byte; 16 readable zero bytes precede it
    uint64_t out_cap;
    uint8_t *win;         // chunk mode: win[p] = window position p
    Node *nodes;          // [256][kNodesStride]  (chain link unused: the decoder keeps no hash chains, lz.rs:348-351)
    uint16_t *words;
    SymRank *symranks;
    uint32_t bhead[kNumBuckets];
    uint32_t spos;        // chunk mode: in/out
    uint32_t after_literal, first_block;
    int32_t self_driven;
    uint64_t out_len;
    uint64_t in_pos;
    int32_t status;
    uint32_t n_chunks;
    unsigned long long dbg[4]; // items, rank-table cache misses
};

// ---- small device helpers --------------------------------------------------------------------
__device__ __forceinline__ uint32_t is_alnum(uint32_t c) { // u8::is_ascii_alphanumeric
    return (((c | 0x20u) - 'a') < 26u) | ((c - '0') < 10u);
}
// matcher.rs:255-263 on a little-endian dword
__device__ __forceinline__ uint32_t hash_dword(uint32_t v) {
    return ((v & 0xffu) * 131313131u ^ 797u) + (((v >> 8) & 0xffu) * 1313131u ^ 79797u) +
           (((v >> 16) & 0xffu) * 13131u ^ 7979797u) + ((v >> 24) * 131u ^ 797979797u);
}
// L2-only load (bypasses L1: coherent with stores of any SM)
__device__ __forceinline__ uint2 ldcg_u2(const uint2 *p) {
#ifdef ORZ_SIMT_EMU
    return *p;
#else
    return __ldcg(p);
#endif
}
// unaligned little-endian dword from a 4-byte-aligned base (global or shared)
__device__ __forceinline__ uint32_t ld_u32_unaligned(const uint8_t *base, uint32_t off) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(base + (off & ~3u));
    return __funnelshift_r(w[0], w[1], (off & 3u) * 8u);
}
// lz.rs:494-534: reduced offset -> (roid, extra bit count, extra bits).  base(id) = 2^(id/2) * (2 + (id&1)) - 2
__device__ __forceinline__ void roid_encode(uint32_t ro, uint32_t &roid, uint32_t &bitlen, uint32_t &bits) {
    uint32_t v = ro + 2;                       // v in [2, 4095]
    uint32_t k = 31u - __clz(v);               // floor(log2 v) >= 1
    uint32_t half = (v >> (k - 1)) & 1u;       // second-highest bit
    roid = 2 * (k - 1) + half;
    bitlen = k - 1;
    bits = v & ((1u << (k - 1)) - 1u);
}
__device__ __forceinline__ uint32_t roid_base(uint32_t roid) { return ((2u + (roid & 1u)) << (roid >> 1)) - 2u; }

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

} // namespace orz
