// decode_kernels.cu — sm_100a kernel of the ROLZ decoder (lz.rs:366-478, lib.rs:94-129).
//
// Decoding one orz stream is a strictly serial chain: the bit position of item n+1 depends on the
// rank-decoded value of item n, which depends on bytes already produced.  One warp owns one stream;
// lane 0 runs the item chain, all 32 lanes build the per-chunk decode tables, initialise the rank
// tables and rebase the ring at block boundaries.  Throughput comes from many streams in flight.
#include "orz_dev.cuh"

#include <cstdlib>

namespace orz {

constexpr int kLutBits = 11;
constexpr int kLutSize = 1 << kLutBits;

struct HuffDec {
    uint16_t lut[kLutSize];          // sym | len << 9 for codes of <= kLutBits bits, 0 otherwise
    uint16_t sorted[kSymStride + 8]; // symbols in (len, sym) order (huffman.rs:130)
    uint16_t first_code[16];
    uint16_t count[16];
    uint16_t offs[16];
    uint8_t lens[kSymStride + 8];
    int n;
    int ok;
};

struct DecShared {
    HuffDec h[3];
    uint16_t bhead[kNumBuckets];
    uint16_t vs[kSymStride + 8];
    uint32_t num_counted;
    int16_t ctag[128];   // rank-table cache tags (context id, -1 = empty): set s holds slots 2s and 2s+1
    uint8_t clru[64];    // per set: the way to replace next (the one not used last)
};

// Two-way set-associative, write-back cache of rank tables in shared memory (few-stream launches: one block per SM).
// Every item reads and writes its context's table (symrank.rs:49-97: three dependent loads); in global memory each of them is an
// L2 round trip, because the item's own stores keep evicting the lines from L1.  A miss swaps 1576 bytes: the fill is issued first
// (independent loads, one L2 latency), the victim is written back while it is in flight.  Called by the whole warp in lock step.
template <int CACHE>
__device__ __forceinline__ SymRank *rank_table(SymRank *symranks, SymRank *cache, DecShared &sh, uint32_t ctx, unsigned &misses) {
    if (CACHE == 0) return &symranks[ctx];
    constexpr uint32_t kSets = CACHE / 2;
    const uint32_t set = ctx % kSets;
    const int t0 = sh.ctag[2 * set], t1 = sh.ctag[2 * set + 1];
    if (t0 == (int)ctx) {
        sh.clru[set] = 1;
        return &cache[2 * set];
    }
    if (t1 == (int)ctx) {
        sh.clru[set] = 0;
        return &cache[2 * set + 1];
    }
    misses++;
    const int lane = threadIdx.x & 31;
    const uint32_t way = sh.clru[set], slot = 2 * set + way;
    const int tag = way ? t1 : t0;
    constexpr int kW = (int)(sizeof(SymRank) / 8); // 197 eight-byte words
    constexpr int kR = (kW + 31) / 32;
    uint2 *c64 = reinterpret_cast<uint2 *>(&cache[slot]);
    const uint2 *s64 = reinterpret_cast<const uint2 *>(&symranks[ctx]);
    uint2 fill[kR];
#pragma unroll
    for (int j = 0; j < kR; j++) {
        const int i = lane + 32 * j;
        fill[j] = i < kW ? ldcg_u2(&s64[i]) : make_uint2(0, 0);
    }
    __syncwarp();
    if (tag >= 0) {
        uint2 *g64 = reinterpret_cast<uint2 *>(&symranks[tag]);
#pragma unroll
        for (int j = 0; j < kR; j++) {
            const int i = lane + 32 * j;
            if (i < kW) g64[i] = c64[i];
        }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kR; j++) {
        const int i = lane + 32 * j;
        if (i < kW) c64[i] = fill[j];
    }
    sh.ctag[slot] = (int16_t)ctx;
    sh.clru[set] = (uint8_t)(way ^ 1u);
    __syncwarp();
    return &cache[slot];
}

struct BitReader { // coder.rs:91-157,159-196
    const uint8_t *base; // 4 B aligned
    uint64_t pos;        // byte offset of the next refill
    uint64_t lim;        // reads are clamped to this offset (end of the input; the allocation is padded past it)
    uint64_t bb;
    int bl;
    __device__ __forceinline__ void refill() {
        if (bl < 32) {
            if (pos > lim) pos = lim; // a malformed header must not walk the reader out of the buffer (coder.rs:152-156 over-reads <= 4 B)
            const uint32_t *wp = reinterpret_cast<const uint32_t *>(base + (pos & ~3ull));
            const uint32_t w = bswap32(__funnelshift_r(wp[0], wp[1], ((uint32_t)pos & 3u) * 8u));
            bb = (bb << 32) | w;
            bl += 32;
            pos += 4;
        }
    }
    __device__ __forceinline__ uint32_t get(int n) {
        refill();
        bl -= n;
        return (uint32_t)(bb >> bl) & ((1u << n) - 1u);
    }
    __device__ uint32_t varint() { // coder.rs:108-120
        uint32_t v = 0;
        for (uint32_t sh = 0;; sh++) {
            const uint32_t b = get(2);
            if (sh < 32) v |= (b & 1u) << sh;
            if (!(b > 1)) break;
            if (sh > 40) break; // malformed
        }
        return v;
    }
};

// coder.rs:127-142 — lane 0 only
__device__ void read_table(BitReader &br, HuffDec &h) {
    const uint32_t maxl = br.varint();
    int n = 0;
    h.ok = maxl <= 15;
    for (;;) {
        const uint32_t delta = br.varint();
        if (delta == 0) break;
        if ((uint32_t)n + delta > (uint32_t)kSymStride) {
            h.ok = 0;
            break;
        }
        for (uint32_t k = 1; k < delta; k++) h.lens[n++] = 0;
        const uint32_t d = br.varint();
        if (d > maxl) h.ok = 0; // the reference would wrap a u8 here; a well-formed table never does
        const uint32_t l = maxl - d;
        h.lens[n++] = (uint8_t)(l & 15u);
    }
    h.n = n;
}

// huffman.rs:119-141,150-167 — canonical codes -> primary LUT + (first_code,count,offs,sorted) for long codes
__device__ void build_table(HuffDec &h, int lane) {
    for (int i = lane; i < kLutSize; i += 32) h.lut[i] = 0;
    if (lane == 0) {
        for (int l = 0; l < 16; l++) h.count[l] = 0;
        for (int s = 0; s < h.n; s++) h.count[h.lens[s]]++;
        h.count[0] = 0;
        uint32_t code = 0, off = 0;
        for (int l = 1; l <= 15; l++) {
            h.first_code[l] = (uint16_t)code;
            h.offs[l] = (uint16_t)off;
            code = (code + h.count[l]) << 1;
            off += h.count[l];
            if (code > (2u << l)) h.ok = 0; // over-subscribed code
        }
        uint16_t fill[16];
        for (int l = 1; l <= 15; l++) fill[l] = h.offs[l];
        for (int s = 0; s < h.n; s++) {
            const int l = h.lens[s];
            if (l) h.sorted[fill[l]++] = (uint16_t)s;
        }
    }
    __syncwarp();
    if (!h.ok) return;
    for (int l = 1; l <= kLutBits; l++) {
        const int cnt = h.count[l];
        const uint32_t fc = h.first_code[l];
        const int span = 1 << (kLutBits - l);
        for (int j = lane; j < cnt; j += 32) {
            const uint16_t sym = h.sorted[h.offs[l] + j];
            const uint32_t basei = (fc + j) << (kLutBits - l);
            const uint16_t e = (uint16_t)(sym | (l << 9));
            for (int q = 0; q < span; q++)
                if (basei + q < (uint32_t)kLutSize) h.lut[basei + q] = e;
        }
    }
    __syncwarp();
}

__device__ __forceinline__ int decode_sym(BitReader &br, const HuffDec &h) { // coder.rs:144-150
    br.refill();
    const uint32_t peek = (uint32_t)(br.bb >> (br.bl - 15)) & 0x7fffu;
    const uint32_t e = h.lut[peek >> (15 - kLutBits)];
    if (e) {
        br.bl -= (int)(e >> 9);
        return (int)(e & 511u);
    }
    for (int l = kLutBits + 1; l <= 15; l++) {
        const uint32_t d = (peek >> (15 - l)) - h.first_code[l];
        if (d < h.count[l]) {
            br.bl -= l;
            return h.sorted[h.offs[l] + d];
        }
    }
    return -1;
}

// symrank.rs:61-97 on a table in global memory
__device__ __forceinline__ void symrank_update(SymRank *s, uint32_t v, uint32_t i, uint32_t ecnt, uint32_t esum) {
    if (ecnt > (uint32_t)kNumSym) {
        ecnt = ecnt * 9 / 10;
        esum = esum * 9 / 10;
    }
    ecnt += 1;
    esum += i;
    s->cnt = ecnt;
    s->sum = esum;
    const uint32_t sub = ((i >> 4) + ((esum >> 4) / ecnt)) & 0xffffu;
    uint32_t next_i = i > sub ? i - sub : 0;
    if (next_i < (i >> 1)) next_i = i >> 1;
    const uint32_t d = i - next_i;
    if (d == 1) {
        const uint16_t nv1 = s->value[next_i];
        s->index[v] = (uint16_t)next_i;
        s->value[i] = nv1;
        s->index[nv1] = (uint16_t)i;
        s->value[next_i] = (uint16_t)v;
    } else if (d > 1) {
        const uint32_t ni2 = next_i, ni1 = next_i + d / 2;
        const uint16_t nv1 = s->value[ni1];
        const uint16_t nv2 = s->value[ni2];
        s->value[i] = nv1;
        s->index[nv1] = (uint16_t)i;
        s->value[ni1] = nv2;
        s->index[nv2] = (uint16_t)ni1;
        s->value[ni2] = (uint16_t)v;
        s->index[v] = (uint16_t)ni2;
    }
}

template <int CACHE>
__global__ void __launch_bounds__(32) orz_decode_kernel(DecStream *streams) {
    DecStream *S = &streams[blockIdx.x];
    if (S->status != 0) return;
    ORZ_DYN_SHARED(smem_raw);
    DecShared &sh = *reinterpret_cast<DecShared *>(smem_raw);
    // on-chip mode (CACHE > 0): the words table (lz.rs:52) and CACHE rank tables live in shared memory
    uint16_t *wsm = reinterpret_cast<uint16_t *>(smem_raw + ((sizeof(DecShared) + 15) & ~size_t(15)));
    SymRank *cache = reinterpret_cast<SymRank *>(smem_raw + ((sizeof(DecShared) + 15) & ~size_t(15)) + (CACHE ? 65536 : 0));
    const int lane = threadIdx.x;

    for (int i = lane; i < kNumBuckets; i += 32) sh.bhead[i] = (uint16_t)S->bhead[i];
    for (int i = lane; i < 128; i += 32) sh.ctag[i] = -1;
    for (int i = lane; i < 64; i += 32) sh.clru[i] = 0;
    Node *nodes = S->nodes;
    if (CACHE)
        for (int i = lane; i < 32768 / 8; i += 32) reinterpret_cast<uint4 *>(wsm)[i] = reinterpret_cast<const uint4 *>(S->words)[i];
    uint16_t *words = CACHE ? wsm : S->words;
    SymRank *symranks = S->symranks;
    const bool self_driven = S->self_driven != 0;
    const uint8_t *in = S->in;
    const uint64_t in_len = S->in_len;
    uint64_t in_pos = S->in_pos;
    uint64_t out_len = S->out_len;       // stream mode: bytes produced so far
    uint64_t blk_apos = 0;               // stream mode: absolute output offset of window position PREMATCH
    uint32_t spos = self_driven ? kPrematch : S->spos;
    uint32_t after_literal = S->after_literal;
    uint32_t first_block = S->first_block;
    uint32_t n_chunks = 0;
    unsigned misses = 0;
    unsigned long long items_done = 0;
    int status = 0;
    __syncwarp();

    for (;;) {
        // ---- framing (lib.rs:108-114, ioutil.rs:60-76) ----
        uint32_t t = 0;
        if (self_driven) {
            if (lane == 0) {
                uint64_t factor = 1, tt = 0;
                for (;;) {
                    if (in_pos >= in_len) {
                        status = -3; // UnexpectedEof
                        break;
                    }
                    const uint32_t v = in[in_pos++];
                    if (v < 128) {
                        tt += v * factor;
                        break;
                    }
                    tt += (v - 128) * factor;
                    factor *= 128;
                    if (factor > (1ull << 42)) {
                        status = -1;
                        break;
                    }
                }
                if (status == 0) {
                    if (tt == 0) status = 1;                      // EOF marker
                    else if (tt >= kTbufLen) status = -1;         // lib.rs:111-113
                    else if (in_pos + tt > in_len) status = -3;   // lib.rs:114
                    t = (uint32_t)tt;
                }
            }
            status = __shfl_sync(0xffffffffu, status, 0);
            if (status != 0) break;
            t = __shfl_sync(0xffffffffu, t, 0);
            in_pos = __shfl_sync(0xffffffffu, in_pos, 0);
        } else {
            t = (uint32_t)in_len;
        }

        // ---- chunk header (lz.rs:372-409) ----
        BitReader br;
        br.base = in;
        br.pos = self_driven ? in_pos : 0;
        br.lim = in_len; // in stream mode in_pos + t <= in_len was checked above
        br.bb = 0;
        br.bl = 0;
        uint32_t sbuf_len = 0, n_items = 0;
        { // every lane parses the header (same loads, same values): no broadcast needed, all lanes stay usable
            if (first_block) {
                uint32_t num = br.varint();
                if (num > (uint32_t)kNumSym) status = -1;
                else {
                    for (int i = 0; i < kSymStride; i++) sh.h[0].lens[i] = 0; // borrow as the `set` array
                    for (uint32_t i = 0; i < num; i++) {
                        const uint32_t v = br.get(9);
                        if (v >= (uint32_t)kNumSym || sh.h[0].lens[v]) {
                            status = -1;
                            break;
                        }
                        sh.vs[i] = (uint16_t)v;
                        sh.h[0].lens[v] = 1;
                    }
                    if (status == 0)
                        for (int i = 0; i < kNumSym; i++)
                            if (!sh.h[0].lens[i]) sh.vs[num++] = (uint16_t)i;
                }
            }
            if (status == 0) {
                sbuf_len = br.varint();
                n_items = br.varint();
                read_table(br, sh.h[0]);
                read_table(br, sh.h[1]);
                read_table(br, sh.h[2]);
                if (!sh.h[0].ok || !sh.h[1].ok || !sh.h[2].ok) status = -1;
                if (n_items > kChunkItems || sbuf_len > kBlockSize) status = -1;
            }
        }
        status = __shfl_sync(0xffffffffu, status, 0);
        if (status != 0) break;
        __syncwarp();
        if (first_block) { // lz.rs:386-391
            for (int tix = 0; tix < kNumCtx; tix++) {
                SymRank *g = &symranks[tix];
                for (int i = lane; i < kNumSym; i += 32) {
                    const uint16_t v = sh.vs[i];
                    g->value[i] = v;
                    g->index[v] = (uint16_t)i;
                }
                if (lane == 0) {
                    g->cnt = 0;
                    g->sum = 1000000u;
                }
            }
            first_block = 0;
        }
        build_table(sh.h[0], lane);
        build_table(sh.h[1], lane);
        build_table(sh.h[2], lane);
        if (!sh.h[0].ok || !sh.h[1].ok || !sh.h[2].ok) {
            status = -1;
            break;
        }
        __threadfence_block();
        __syncwarp();

        // ---- items (lz.rs:411-476): the serial chain, lane 0 ----
        const uint32_t spos_start = spos;
        { // the serial chain, executed by every lane in lock step: identical loads, identical (redundant) stores, so
          // each lane sees its own writes and all 32 lanes are at hand for the cooperative copies
            // window position p lives at win[p]; `wa`/`dl` give a 16 B aligned base for dword accesses
            // (stream mode decodes in place: position PREMATCH of the current block is out[blk_apos]).
            uint8_t *wa = self_driven ? (S->out + blk_apos - kBlockNew) : S->win;
            const uint32_t dl = self_driven ? 1u : 0u;
            uint8_t *win = wa + dl;
            const uint64_t cap_pos = self_driven ? (uint64_t)kPrematch + (S->out_cap - blk_apos) : (uint64_t)kBlockSize;
            const uint32_t lo_pos = self_driven ? (blk_apos >= kPrematch ? 0u : (uint32_t)(kPrematch - blk_apos)) : 0u;
            const uint32_t limit = (uint32_t)(cap_pos < (uint64_t)kBlockSize ? cap_pos : (uint64_t)kBlockSize);
            // bytes spos-4..spos-1, little endian (byte 3 = spos-1)
            uint32_t prev4 = ld_u32_unaligned(wa, spos - 4 + dl);
            for (uint32_t k = 0; k < n_items; k++) {
                const uint32_t b_m1 = prev4 >> 24, b_m2 = (prev4 >> 16) & 0xffu, b_m3 = (prev4 >> 8) & 0xffu;
                const uint32_t c = (b_m1 & 0x7fu) | is_alnum(b_m2) << 7;
                const uint32_t h2 = (b_m1 & 0x7fu) | ((b_m2 & 0x7fu) | is_alnum(b_m3) << 7) << 7;
                const uint32_t lwe = words[h2];
                SymRank *sr = rank_table<CACHE>(symranks, cache, sh, c | after_literal << 8, misses);
                const uint32_t ecnt0 = sr->cnt, esum0 = sr->sum; // issued early: independent of the bit stream
                const uint32_t iun = sr->index[lwe & 0xffu];
                const int sym = decode_sym(br, sh.h[after_literal]);
                if (sym < 0) {
                    status = -1;
                    break;
                }
                const uint32_t i = (sym == kNumSym - 1) ? iun : (uint32_t)sym + ((uint32_t)sym >= iun); // symrank.rs:49-55
                if (i >= (uint32_t)kNumSym) {
                    status = -1;
                    break;
                }
                const uint32_t v = sr->value[i];
                symrank_update(sr, v, i, ecnt0, esum0);

                Node *bn = nodes + c * kNodesStride;
                const uint32_t head = sh.bhead[c];
                const uint32_t new_head = head + 1 == kRing ? 0 : head + 1;
                Node nn;
                nn.pl = spos & 0x1ffffffu;
                nn.next = -1;
                nn.pad = 0;
                if (v == (uint32_t)kWordSym) { // lz.rs:425-430
                    if (spos + 2 > limit) {
                        status = -2;
                        break;
                    }
                    nn.expected = 0;
                    bn[new_head] = nn;
                    sh.bhead[c] = (uint16_t)new_head;
                    after_literal = 0;
                    win[spos] = (uint8_t)(lwe & 0xffu);
                    win[spos + 1] = (uint8_t)(lwe >> 8);
                    spos += 2;
                    prev4 = (prev4 >> 16) | lwe << 16;
                } else if (v < 256) { // lz.rs:431-437
                    if (spos + 1 > limit) {
                        status = -2;
                        break;
                    }
                    nn.expected = 0;
                    bn[new_head] = nn;
                    sh.bhead[c] = (uint16_t)new_head;
                    after_literal = 1;
                    win[spos] = (uint8_t)v;
                    spos += 1;
                    const uint32_t y5 = prev4 & 0xffu; // byte spos'-5
                    prev4 = (prev4 >> 8) | v << 24;    // bytes spos'-4 .. spos'-1
                    const uint32_t y4 = prev4 & 0xffu, y3 = (prev4 >> 8) & 0xffu;
                    const uint32_t hw = (y3 & 0x7fu) | ((y4 & 0x7fu) | is_alnum(y5) << 7) << 7; // hash2(spos'-3)
                    words[hw] = (uint16_t)(prev4 >> 16);
                } else { // lz.rs:438-474
                    const uint32_t roid = (v - 256) / kLenidSize, lenid = (v - 256) % kLenidSize;
                    const uint32_t robitlen = roid >> 1;
                    const uint32_t ro = roid_base(roid) + (robitlen ? br.get((int)robitlen) : 0u);
                    if (ro >= (uint32_t)kRing) {
                        status = -1;
                        break;
                    }
                    const uint32_t node = (head + kRing - ro) % kRing;
                    const unsigned long long nrec = *reinterpret_cast<const unsigned long long *>(&bn[node]);
                    const uint32_t pl = (uint32_t)nrec;
                    uint32_t mexp = (uint32_t)(nrec >> 48) & 0xffu;
                    uint32_t eml = lenid;
                    if (lenid == (uint32_t)(kLenidSize - 1)) {
                        const int e2 = decode_sym(br, sh.h[2]);
                        if (e2 < 0) {
                            status = -1;
                            break;
                        }
                        eml = (uint32_t)e2;
                    }
                    const uint32_t mpos = node_pos(pl);
                    uint32_t mmin = node_lenmin(pl);
                    if (mmin < (uint32_t)kMatchMin) mmin = kMatchMin;
                    if (mexp < (uint32_t)kMatchMin) mexp = kMatchMin;
                    uint32_t mlen;
                    if (eml + mmin > mexp) mlen = eml + mmin;
                    else if (eml > 0) mlen = eml + mmin - 1;
                    else mlen = mexp;
                    if (mpos >= spos || mpos < lo_pos || mlen > 0x1ffu) {
                        status = -1;
                        break;
                    }
                    if (spos + mlen > limit) {
                        status = -2;
                        break;
                    }
                    // Bucket::update (matcher.rs:62-80)
                    if (node_lenmin(pl) <= mlen) {
                        const uint32_t lm = mlen + 1 < 127u ? mlen + 1 : 127u;
                        bn[node].pl = mpos | lm << 25;
                    }
                    nn.expected = (uint8_t)mlen;
                    bn[new_head] = nn;
                    sh.bhead[c] = (uint16_t)new_head;
                    after_literal = 0;
                    // mem.rs:74-92: semantically a byte-serial overlapped copy.  The tail bytes needed for the next
                    // hashes are carried in registers: a load right after a store to the same line misses L1.
                    const uint32_t dist = spos - mpos;
                    uint32_t y5;
                    if (dist >= mlen && mlen <= 256u) {
                        // source and destination do not overlap: the whole copy in one round trip, lane l moves bytes 8l .. 8l+7
                        const uint32_t off = 8u * (uint32_t)lane;
                        uint32_t lo = 0, hi = 0;
                        if (off < mlen) {
                            lo = ld_u32_unaligned(wa, mpos + off + dl);
                            if (off + 4 < mlen) hi = ld_u32_unaligned(wa, mpos + off + 4 + dl);
                            const uint32_t nb = mlen - off; // bytes of this lane (>= 1)
                            uint8_t *d = win + spos + off;
                            d[0] = (uint8_t)lo;
                            if (nb > 1) d[1] = (uint8_t)(lo >> 8);
                            if (nb > 2) d[2] = (uint8_t)(lo >> 16);
                            if (nb > 3) d[3] = (uint8_t)(lo >> 24);
                            if (nb > 4) d[4] = (uint8_t)hi;
                            if (nb > 5) d[5] = (uint8_t)(hi >> 8);
                            if (nb > 6) d[6] = (uint8_t)(hi >> 16);
                            if (nb > 7) d[7] = (uint8_t)(hi >> 24);
                        }
                        // the five bytes before the new position (for the next hashes): they sit in lane A and the lane before it
                        // (or in prev4 when A == 0); positions 8A-8 .. 8A+7 of the match as one 128-bit value, bytes idx-4 .. idx wanted
                        const uint32_t A = (mlen - 1) >> 3;
                        const uint32_t alo = __shfl_sync(0xffffffffu, lo, (int)A), ahi = __shfl_sync(0xffffffffu, hi, (int)A);
                        const uint32_t plo = __shfl_sync(0xffffffffu, lo, (int)(A ? A - 1 : 0)), phi = __shfl_sync(0xffffffffu, hi, (int)(A ? A - 1 : 0));
                        const unsigned long long vA = (unsigned long long)ahi << 32 | alo;
                        const unsigned long long vP = A ? ((unsigned long long)phi << 32 | plo) : ((unsigned long long)prev4 << 32);
                        const uint32_t sh8 = (mlen - 1 - 8u * A + 4u) * 8u; // (idx - 4) * 8 with idx = mlen - 1 - 8A + 8: 32 .. 88
                        const unsigned long long x5 = sh8 < 64u ? (vP >> sh8) | (vA << (64u - sh8)) : (vA >> (sh8 - 64u));
                        y5 = (uint32_t)x5 & 0xffu;
                        prev4 = (uint32_t)(x5 >> 8);
                    } else if (dist >= 4) {
                        uint32_t lo = prev4, hi = 0, left = 0;
                        for (uint32_t q = 0; q < mlen; q += 4) {
                            const uint32_t wv = ld_u32_unaligned(wa, mpos + q + dl);
                            left = mlen - q;
                            win[spos + q] = (uint8_t)wv;
                            if (left > 1) win[spos + q + 1] = (uint8_t)(wv >> 8);
                            if (left > 2) win[spos + q + 2] = (uint8_t)(wv >> 16);
                            if (left > 3) win[spos + q + 3] = (uint8_t)(wv >> 24);
                            if (q) lo = hi;
                            hi = wv;
                        }
                        if (left > 4) left = 4; // the last iteration wrote `left` bytes of hi after the 4 bytes of lo
                        const unsigned long long both = (unsigned long long)hi << 32 | lo;
                        y5 = (uint32_t)(both >> (8 * (left - 1))) & 0xffu; // byte end-5
                        prev4 = (uint32_t)(both >> (8 * left));            // bytes end-4 .. end-1
                    } else { // period-dist run: replicate the last `dist` bytes from registers
                        uint32_t pat = prev4 >> (8 * (4 - dist)); // the last dist bytes, lowest address first
                        y5 = 0;
                        for (uint32_t q = 0; q < mlen; q++) {
                            const uint32_t byte = pat & 0xffu;
                            win[spos + q] = (uint8_t)byte;
                            y5 = prev4 & 0xffu;
                            prev4 = (prev4 >> 8) | byte << 24;
                            pat = (pat >> 8) | byte << (8 * (dist - 1)); // rotate within dist bytes
                        }
                        // y5 must be byte end-5: prev4 before the last shift held end-5 in its lowest byte
                    }
                    spos += mlen;
                    const uint32_t y4 = prev4 & 0xffu, y3 = (prev4 >> 8) & 0xffu;
                    const uint32_t hw = (y3 & 0x7fu) | ((y4 & 0x7fu) | is_alnum(y5) << 7) << 7; // hash2(spos-3)
                    words[hw] = (uint16_t)(prev4 >> 16);
                }
            }
            if (status == 0 && spos != sbuf_len) {
                // lz.rs:477 returns min(spos, sbuf_len); a well-formed chunk ends exactly at sbuf_len
                if (spos > sbuf_len) spos = sbuf_len;
                else status = -1;
            }
        }
        status = __shfl_sync(0xffffffffu, status, 0);
        const uint32_t spos_end = __shfl_sync(0xffffffffu, spos, 0);
        after_literal = __shfl_sync(0xffffffffu, after_literal, 0);
        n_chunks++;
        items_done += n_items;
        if (status != 0) break;
        __threadfence_block();
        __syncwarp();
        if (!self_driven) {
            spos = spos_end;
            break;
        }
        out_len += spos_end - spos_start; // lib.rs:117
        spos = spos_end;
        in_pos += t;
        if (spos >= kBlockSize) { // lib.rs:120-125: slide (a pointer shift here) + LZDecoder::forward (lz.rs:359-364)
            if (spos > kBlockSize) {
                status = -1;
                break;
            }
            blk_apos += kBlockNew;
            spos = kPrematch;
            for (int i = lane; i < kNumBuckets * kNodesStride; i += 32) {
                if ((i & (kNodesStride - 1)) >= kRing) continue;
                const uint32_t pl = nodes[i].pl;
                const uint32_t p = node_pos(pl);
                nodes[i].pl = (p > kBlockNew ? p - kBlockNew : 0u) | (pl & 0xfe000000u);
            }
            __threadfence_block();
            __syncwarp();
        }
    }

    __syncwarp();
    if (CACHE > 0) { // write the words table and the cached rank tables back (the state persists across chunk-mode calls)
        for (int i = lane; i < 32768 / 8; i += 32) reinterpret_cast<uint4 *>(S->words)[i] = reinterpret_cast<const uint4 *>(wsm)[i];
        for (int sl = 0; sl < CACHE; sl++) {
            const int tag = sh.ctag[sl];
            if (tag < 0) continue;
            const uint32_t *c32 = reinterpret_cast<const uint32_t *>(&cache[sl]);
            uint32_t *g32 = reinterpret_cast<uint32_t *>(&symranks[tag]);
            for (int i = lane; i < (int)(sizeof(SymRank) / 4); i += 32) g32[i] = c32[i];
        }
    }
    __syncwarp();
    for (int i = lane; i < kNumBuckets; i += 32) S->bhead[i] = sh.bhead[i];
    if (lane == 0) {
        S->spos = spos;
        S->after_literal = after_literal;
        S->first_block = first_block;
        S->out_len = out_len;
        S->in_pos = in_pos;
        S->n_chunks += n_chunks;
        S->dbg[0] += items_done;
        S->dbg[1] += misses;
        S->status = self_driven ? status : (status == 0 ? 0 : status);
    }
}

// LZDecoder::forward for the chunk-level API (lz.rs:359-364) + device-side window slide (lib.rs:121)
__global__ void __launch_bounds__(256) orz_dec_forward_kernel(DecStream *streams, uint32_t forward_len) {
    DecStream *S = &streams[0];
    Node *nodes = S->nodes;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    for (int i = gid; i < kNumBuckets * kNodesStride; i += gsz) {
        if ((i & (kNodesStride - 1)) >= kRing) continue;
        const uint32_t pl = nodes[i].pl;
        const uint32_t p = node_pos(pl);
        nodes[i].pl = (p > forward_len ? p - forward_len : 0u) | (pl & 0xfe000000u);
    }
}

constexpr int kDecCache = 94; // rank tables cached per stream in on-chip mode: 47 sets x 2 ways beside the 64 KB words table
void launch_decode(DecStream *d_streams, int n, cudaStream_t st) {
    constexpr size_t base = (sizeof(DecShared) + 15) & ~size_t(15);
    constexpr size_t big = base + 65536 + kDecCache * sizeof(SymRank);
    static_assert(big <= 232448, "decoder: shared memory of the on-chip mode");
    // (attributes are per device: set on every launch, it is a cheap call)
    cudaFuncSetAttribute(orz_decode_kernel<kDecCache>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big);
    cudaFuncSetAttribute(orz_decode_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)base);
    // few streams (at most two per SM's worth of waves): one block per SM with the words table and 94 rank tables in shared memory,
    // the item chain runs out of on-chip memory.  Many streams: several blocks per SM hide each other's L2 latency instead.
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 2 * sms) ORZ_LAUNCH(orz_decode_kernel<kDecCache>, n, 32, big, st, d_streams);
    else ORZ_LAUNCH(orz_decode_kernel<0>, n, 32, base, st, d_streams);
}
void launch_dec_forward(DecStream *d_stream, uint32_t forward_len, cudaStream_t st) {
    ORZ_LAUNCH(orz_dec_forward_kernel, 148, 256, 0, st, d_stream, forward_len);
}

} // namespace orz
