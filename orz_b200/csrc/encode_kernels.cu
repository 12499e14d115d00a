// encode_kernels.cu — sm_100a kernels of the ROLZ encoder.
//
//   orz_parse_kernel   stage A  (lz.rs:131-235): one warp owns one stream; the ROLZ parse is a serial
//                      state machine, the lanes cooperate inside a step (parallel candidate probes + LCP,
//                      ballot-resolved sequential semantics of matcher.rs:135-192).
//   orz_emit_kernel    stages B-D (lz.rs:238-344): one 512-thread block per stream — per-context rank
//                      coding (one warp per context, table in shared memory), shared-memory histograms,
//                      length-limited canonical Huffman build, block-wide prefix scan of bit lengths and
//                      big-endian bit-pack, framing (lib.rs:79-80) and the block loop (lib.rs:72-86).
//
// Integer/byte work, latency bound: no tensor cores.  All reference citations are file:line into
// richox/orz @29414c6.
#include "orz_dev.cuh"

namespace orz {

// ================================================================================================
// Stage A
// ================================================================================================
constexpr int kTile = 2048;      // bytes of the window staged in shared memory around spos
constexpr int kTileNeed = 432;   // bytes that must be staged at and after spos (240 LCP + lazy + slack)
constexpr int kParseThreads = 96; // warp 0: find_match at spos; warp 1 / 2: lazy probes at spos+1 / spos+2
constexpr unsigned kFull = 0xffffffffu;

// Reads the head record of a chain: pl, expected, and (lane L in 1..kSnap) the snapshot link next[L-1].
__device__ __forceinline__ void enode_head(const ENode *bn, int slot, uint32_t &pl, uint32_t &exp, int &my_next, int lane) {
    const uint32_t *r = reinterpret_cast<const uint32_t *>(bn + slot);
    pl = r[0];
    exp = r[1] & 0xffu;
    my_next = -1;
    if (lane >= 1 && lane <= kSnap) {
        const uint32_t w = r[(lane + 2) >> 1]; // int16 index lane+2 of the record
        my_next = (int)(int16_t)(((lane + 2) & 1) ? (w >> 16) : (w & 0xffffu));
    }
}

// Longest common prefix of window[npos..] and the staged bytes tile[toff..], capped at 240
// (mem.rs:41-51: 16-byte steps, same result as a byte-wise LCP capped at max_len).
__device__ __forceinline__ int lcp_win_tile(const uint8_t *win, uint32_t npos, const uint8_t *tile, uint32_t toff) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(win + (npos & ~3u));
    const uint32_t sh = (npos & 3u) * 8u;
    const uint32_t *t = reinterpret_cast<const uint32_t *>(tile + (toff & ~3u));
    const uint32_t tsh = (toff & 3u) * 8u;
    uint32_t wprev = w[0], tprev = t[0];
    for (int i = 0; i < kMatchMax / 16; i++) {
        uint32_t w1 = w[4 * i + 1], w2 = w[4 * i + 2], w3 = w[4 * i + 3], w4 = w[4 * i + 4];
        uint32_t t1 = t[4 * i + 1], t2 = t[4 * i + 2], t3 = t[4 * i + 3], t4 = t[4 * i + 4];
        uint32_t x0 = __funnelshift_r(wprev, w1, sh) ^ __funnelshift_r(tprev, t1, tsh);
        uint32_t x1 = __funnelshift_r(w1, w2, sh) ^ __funnelshift_r(t1, t2, tsh);
        uint32_t x2 = __funnelshift_r(w2, w3, sh) ^ __funnelshift_r(t2, t3, tsh);
        uint32_t x3 = __funnelshift_r(w3, w4, sh) ^ __funnelshift_r(t3, t4, tsh);
        if (x0) return 16 * i + ((__ffs(x0) - 1) >> 3);
        if (x1) return 16 * i + 4 + ((__ffs(x1) - 1) >> 3);
        if (x2) return 16 * i + 8 + ((__ffs(x2) - 1) >> 3);
        if (x3) return 16 * i + 12 + ((__ffs(x3) - 1) >> 3);
        wprev = w4;
        tprev = t4;
    }
    return kMatchMax;
}

// One gather round of a chain walk.  Lane 0 holds the round's head node X (x_*), lanes 1..cap hold
// X.next[0..cap-1] in my_next.  Loads those records in parallel and validates the chain order
// (matcher.rs:171-179: link != -1, positions strictly decrease).  Returns the first invalid lane.
__device__ __forceinline__ int gather_round(const ENode *bn, int lane, int cap, uint32_t x_pl, uint32_t x_exp, int x_idx,
                                            int my_next, uint32_t &my_pl, uint32_t &my_exp, int &my_idx) {
    bool ok = false;
    my_pl = 0;
    my_exp = 0;
    my_idx = -1;
    if (lane == 0) {
        my_pl = x_pl;
        my_exp = x_exp;
        my_idx = x_idx;
        ok = true;
    } else if (lane <= cap && my_next >= 0) {
        const uint32_t *rr = reinterpret_cast<const uint32_t *>(bn + my_next);
        my_pl = rr[0];
        my_exp = rr[1] & 0xffu;
        my_idx = my_next;
        ok = true;
    }
    const uint32_t prev_pos = __shfl_up_sync(kFull, node_pos(my_pl), 1);
    const bool link_ok = (lane == 0) || (ok && node_pos(my_pl) < prev_pos);
    const unsigned bad = ~__ballot_sync(kFull, link_ok);
    return bad ? (__ffs(bad) - 1) : 32;
}

struct FoundMatch {
    int len;          // 0 = none
    int idx;          // best node slot
    uint32_t pl;      // best node pos|lenmin (raw)
    uint32_t exp;     // best node expected (raw)
};

// matcher.rs:135-192.  h0 = heads[entry].  On return lane L (1..kSnap) holds in f_next the head node's snapshot
// link next[L-1] (used to build the new slot's own snapshot), -1 if there was no head.
__device__ __forceinline__ FoundMatch find_match_warp(const uint8_t *win, const ENode *bn, int h0, const uint8_t *tile,
                                                      uint32_t toff, uint32_t pos, uint32_t buf_len, int depth, int lane,
                                                      int &f_next) {
    FoundMatch r;
    r.len = 0;
    r.idx = 0;
    r.pl = 0;
    r.exp = 0;
    f_next = -1;
    if (h0 < 0) return r;
    uint32_t x_pl, x_exp;
    int x_idx = h0, my_next;
    enode_head(bn, h0, x_pl, x_exp, my_next, lane);
    f_next = my_next;
    if (depth <= 0) return r;

    int max_len = kMatchMin - 1;
    uint32_t exp_best = kMatchMin, pl_best = (uint32_t)kMatchMin << 25;
    int idx_best = 0;
    int remaining = depth;
    bool done = false;
    bool first_round = true;

    for (;;) {
        const int lo = first_round ? 0 : 1; // the head node is a candidate only in the first round
        const int room = remaining - (first_round ? 1 : 0);
        const int cap = room < kSnap ? room : kSnap;
        uint32_t my_pl, my_exp;
        int my_idx;
        const int nvl = gather_round(bn, lane, cap, x_pl, x_exp, x_idx, my_next, my_pl, my_exp, my_idx);
        const bool valid = lane >= lo && lane < nvl;
        int lcp = -1;
        if (valid) lcp = lcp_win_tile(win, node_pos(my_pl), tile, toff);

        // Resolve the sequential semantics of matcher.rs:149-181 over this round's candidates.
        int cur = lo - 1;
        for (;;) {
            const unsigned after = (cur >= 31) ? 0u : (0xffffffffu << (cur + 1));
            const unsigned upd = __ballot_sync(kFull, valid && lcp > max_len) & after;
            const int u = upd ? (__ffs(upd) - 1) : 32;
            // candidates whose probe dword (matcher.rs:150-153) may match although they are not longer than
            // max_len: they cannot update, but matcher.rs:165-167 may still break on them.
            unsigned brk = __ballot_sync(kFull, valid && exp_best > 0 && lcp < max_len - 3 && lcp > (int)exp_best) & after;
            if (u < 32) brk &= (1u << u) - 1u;
            if (brk) {
                bool eq = false;
                if ((brk >> lane) & 1u) {
                    const uint32_t a = ld_u32_unaligned(win, node_pos(my_pl) + max_len - 3);
                    const uint32_t b = ld_u32_unaligned(tile, toff + max_len - 3);
                    eq = (a == b);
                }
                if (__ballot_sync(kFull, eq)) {
                    done = true;
                    break;
                }
            }
            if (u >= 32) break;
            max_len = __shfl_sync(kFull, lcp, u);
            exp_best = __shfl_sync(kFull, my_exp, u);
            pl_best = __shfl_sync(kFull, my_pl, u);
            idx_best = __shfl_sync(kFull, my_idx, u);
            if (max_len == kMatchMax || (exp_best > 0 && max_len > (int)exp_best)) { // matcher.rs:162-167
                done = true;
                break;
            }
            cur = u;
        }
        remaining -= nvl - lo;
        if (done || remaining <= 0 || nvl <= cap) break; // nvl <= cap: the chain ended inside this round
        // continue from the last node of this round (lane cap == kSnap): it becomes the next round's head
        x_pl = __shfl_sync(kFull, my_pl, kSnap);
        x_exp = __shfl_sync(kFull, my_exp, kSnap);
        x_idx = __shfl_sync(kFull, my_idx, kSnap);
        uint32_t d0, d1;
        enode_head(bn, x_idx, d0, d1, my_next, lane);
        first_round = false;
    }
    if (max_len >= kMatchMin && pos + max_len < buf_len) { // matcher.rs:183
        r.len = max_len;
        r.idx = idx_best;
        r.pl = pl_best;
        r.exp = exp_best;
    }
    return r;
}

// matcher.rs:194-228 restated as "the longest LCP among the first `depth` chain nodes": has_lazy_match(min_len)
// is true iff that maximum is >= min_len, so the probe can run before min_len is known.  Returns -1 for no node.
__device__ __forceinline__ int lazy_max_lcp_warp(const uint8_t *win, const ENode *bn, int h0, const uint8_t *tile,
                                                 uint32_t toff, int depth, int lane) {
    if (h0 < 0 || depth <= 0) return -1;
    uint32_t x_pl, x_exp;
    int x_idx = h0, my_next;
    enode_head(bn, h0, x_pl, x_exp, my_next, lane);
    int remaining = depth;
    bool first_round = true;
    int best = -1;
    for (;;) {
        const int lo = first_round ? 0 : 1;
        const int room = remaining - (first_round ? 1 : 0);
        const int cap = room < kSnap ? room : kSnap;
        uint32_t my_pl, my_exp;
        int my_idx;
        const int nvl = gather_round(bn, lane, cap, x_pl, x_exp, x_idx, my_next, my_pl, my_exp, my_idx);
        int lcp = -1;
        if (lane >= lo && lane < nvl) lcp = lcp_win_tile(win, node_pos(my_pl), tile, toff);
        for (int o = 16; o; o >>= 1) {
            const int other = __shfl_xor_sync(kFull, lcp, o);
            lcp = other > lcp ? other : lcp;
        }
        best = lcp > best ? lcp : best;
        remaining -= nvl - lo;
        if (remaining <= 0 || nvl <= cap || best >= kMatchMax / 2 + 2) break; // lazy lengths never exceed 121
        x_pl = __shfl_sync(kFull, my_pl, kSnap);
        x_exp = __shfl_sync(kFull, my_exp, kSnap);
        x_idx = __shfl_sync(kFull, my_idx, kSnap);
        uint32_t d0, d1;
        enode_head(bn, x_idx, d0, d1, my_next, lane);
        first_round = false;
    }
    return best;
}

struct ParseXchg {
    int fm_len, fm_idx;
    uint32_t fm_pl, fm_exp;
    int lazy_max[2];
};

__global__ void __launch_bounds__(kParseThreads) orz_parse_kernel(EncStream *streams) {
    EncStream *S = &streams[blockIdx.x];
    if (S->status != 0) return;

    __shared__ uint32_t s_cnt[kNumCtx];
    __shared__ uint16_t s_bhead[kNumBuckets];
    __shared__ __align__(16) uint8_t s_tile[kTile + 32];
    __shared__ ParseXchg s_x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < kNumCtx; i += kParseThreads) s_cnt[i] = 0;
    for (int i = tid; i < kNumBuckets; i += kParseThreads) s_bhead[i] = (uint16_t)S->bhead[i];

    const uint8_t *win = S->win;
    ENode *nodes = S->nodes;
    int16_t *heads = S->heads;
    uint16_t *words = S->words;
    Item *items = S->items;
    const uint32_t sbuf_len = S->sbuf_len;
    const uint32_t items_cap = S->items_cap < kChunkItems ? S->items_cap : kChunkItems;
    const int depth0 = (int)S->depth0, depth1 = (int)S->depth1, depth2 = (int)S->depth2;
    uint32_t spos = S->spos;
    uint32_t after_literal = S->after_literal;
    uint32_t n_items = 0;
    uint32_t tbase = 0xffffffffu;
    __syncthreads();

    while (spos < sbuf_len && n_items < items_cap) { // lz.rs:131
        // ---- stage the look-ahead tile ----
        if (tbase == 0xffffffffu || spos + kTileNeed > tbase + kTile) {
            tbase = (spos - 16u) & ~15u;
            for (int i = tid; i < kTile / 16; i += kParseThreads)
                reinterpret_cast<uint4 *>(s_tile)[i] = *reinterpret_cast<const uint4 *>(win + tbase + 16u * i);
            if (tid < 8) reinterpret_cast<uint32_t *>(s_tile + kTile)[tid] = 0;
            __syncthreads();
        }
        const uint32_t to = spos - tbase;
        const uint32_t prev4 = ld_u32_unaligned(s_tile, to - 4); // bytes spos-4 .. spos-1
        const uint32_t cur4 = ld_u32_unaligned(s_tile, to);      // bytes spos .. spos+3
        const uint32_t b_m1 = prev4 >> 24, b_m2 = (prev4 >> 16) & 0xffu, b_m3 = (prev4 >> 8) & 0xffu;
        const uint32_t b0 = cur4 & 0xffu, b1 = (cur4 >> 8) & 0xffu;
        const uint32_t c = (b_m1 & 0x7fu) | is_alnum(b_m2) << 7;                                 // hash1(spos-1)  lz.rs:481-486
        const uint32_t e0 = hash_dword(cur4) % kHashSize;
        int16_t *hp = heads + c * kHeadsStride + e0;
        ENode *bn = nodes + c * kNodesStride;
        int h0 = -1, f_next = -1;
        const uint32_t h2 = (b_m1 & 0x7fu) | ((b_m2 & 0x7fu) | is_alnum(b_m3) << 7) << 7;        // hash2(spos-1)  lz.rs:488-492
        const uint32_t lwe = words[h2];                                                           // lz.rs:132
        const bool last_word_matched = (cur4 & 0xffffu) == lwe;                                   // lz.rs:133

        if (warp == 0) {
            h0 = *hp;
            const FoundMatch m = find_match_warp(win, bn, h0, s_tile, to, spos, sbuf_len, depth0, lane, f_next); // lz.rs:140-145
            if (lane == 0) {
                s_x.fm_len = m.len;
                s_x.fm_idx = m.idx;
                s_x.fm_pl = m.pl;
                s_x.fm_exp = m.exp;
            }
        } else if (warp == 1) { // lz.rs:151-171 probe 1: bucket hash1(spos), position spos+1
            const uint32_t c1 = (b0 & 0x7fu) | is_alnum(b_m1) << 7;
            const uint32_t e1 = hash_dword(ld_u32_unaligned(s_tile, to + 1)) % kHashSize;
            const int h1 = heads[c1 * kHeadsStride + e1];
            const int mx = lazy_max_lcp_warp(win, nodes + c1 * kNodesStride, h1, s_tile, to + 1, depth1, lane);
            if (lane == 0) s_x.lazy_max[0] = mx;
        } else { // probe 2: bucket hash1(spos+1), position spos+2
            const uint32_t c2 = (b1 & 0x7fu) | is_alnum(b0) << 7;
            const uint32_t e2 = hash_dword(ld_u32_unaligned(s_tile, to + 2)) % kHashSize;
            const int h2i = heads[c2 * kHeadsStride + e2];
            const int mx = lazy_max_lcp_warp(win, nodes + c2 * kNodesStride, h2i, s_tile, to + 2, depth2, lane);
            if (lane == 0) s_x.lazy_max[1] = mx;
        }
        __syncthreads();

        // ---- decision (every thread, uniform) ----
        FoundMatch m;
        m.len = s_x.fm_len;
        m.idx = s_x.fm_idx;
        m.pl = s_x.fm_pl;
        m.exp = s_x.fm_exp;
        const uint32_t head = s_bhead[c];
        const uint32_t new_head = head + 1 == kRing ? 0 : head + 1;
        const uint32_t ctx = c | after_literal << 8;                                              // lz.rs:134-135
        int lazy_id = 0;
        uint32_t ro = 0, roid = 0, robitlen = 0, robits = 0;
        if (m.len > 0) {
            ro = (head + kRing - (uint32_t)m.idx) % kRing; // matcher.rs:185
            roid_encode(ro, roid, robitlen, robits);       // lz.rs:148
        }
        if (m.len > 0 && m.len < kMatchMax / 2) {          // lz.rs:151-171
            const int lazy_len1 = m.len + 1 + (robitlen < 8);
            const int lazy_len2 = lazy_len1 - (int)last_word_matched;
            if (s_x.lazy_max[0] >= lazy_len1) lazy_id = 1;
            else if (s_x.lazy_max[1] >= lazy_len2) lazy_id = 2;
        }

        const bool emit_match = m.len > 0 && lazy_id == 0;
        const uint32_t mlen = emit_match ? (uint32_t)m.len : 0u;
        const uint32_t seq = s_cnt[ctx];
        uint32_t new_spos;
        Item it;
        it.ctx = (uint16_t)ctx;
        it.unlikely = (uint8_t)(lwe & 0xffu);                                                     // lz.rs:136
        it.pad = 0;
        it.seq = seq;
        if (emit_match) { // lz.rs:173-209
            const uint32_t mexp = m.exp > (uint32_t)kMatchMin ? m.exp : (uint32_t)kMatchMin;
            const uint32_t mmin = node_lenmin(m.pl) > (uint32_t)kMatchMin ? node_lenmin(m.pl) : (uint32_t)kMatchMin;
            uint32_t eml;
            if (mlen > mexp) eml = mlen - mmin;
            else if (mlen < mexp) eml = mlen - mmin + 1;
            else eml = 0;
            eml &= 0xffu; // `as u8`
            const uint32_t lenid = eml < (uint32_t)(kLenidSize - 1) ? eml : (uint32_t)(kLenidSize - 1);
            it.symbol = (uint16_t)(256 + roid * kLenidSize + lenid);
            it.robits = (uint16_t)robits;
            it.robitlen = (uint8_t)robitlen;
            it.eml = (uint8_t)eml;
            it.flags = (uint8_t)(after_literal | 2u);
            new_spos = spos + mlen;
        } else {
            it.robits = 0;
            it.robitlen = 0;
            it.eml = 0;
            it.flags = (uint8_t)after_literal;
            if (spos + 1 < sbuf_len && lazy_id != 1 && last_word_matched) { // lz.rs:218-226
                it.symbol = (uint16_t)kWordSym;
                new_spos = spos + 2;
            } else { // lz.rs:227-236
                it.symbol = (uint16_t)b0;
                new_spos = spos + 1;
            }
        }

        if (warp == 0) {
            // Bucket::update (matcher.rs:62-80) + BucketMatcher::update (matcher.rs:115-121)
            if (lane == 0 && emit_match && node_lenmin(m.pl) <= mlen && (uint32_t)m.idx != new_head) {
                const uint32_t lm = mlen + 1 < 127u ? mlen + 1 : 127u;
                bn[m.idx].pl = node_pos(m.pl) | lm << 25;
            }
            // the new slot: pl, expected, next[0] = heads[entry] (matcher.rs:119), next[1..] = the old head's snapshot
            const int shifted = __shfl_sync(kFull, f_next, lane >= 4 ? lane - 3 : 0);
            const uint32_t npl = spos & 0x1ffffffu;
            int16_t hv;
            if (lane == 0) hv = (int16_t)(npl & 0xffffu);
            else if (lane == 1) hv = (int16_t)(npl >> 16);
            else if (lane == 2) hv = (int16_t)(mlen & 0xffu);
            else if (lane == 3) hv = (int16_t)h0;
            else hv = (int16_t)shifted;
            reinterpret_cast<int16_t *>(bn + new_head)[lane] = hv;
            if (lane == 0) {
                *hp = (int16_t)new_head;
                s_bhead[c] = (uint16_t)new_head;
                s_cnt[ctx] = seq + 1;
                items[n_items] = it;
                if (it.symbol != kWordSym) { // lz.rs:208,233: words[hash2(spos'-3)] = buf[spos'-2..spos']
                    const uint32_t q = new_spos - tbase;
                    const uint32_t t4 = ld_u32_unaligned(s_tile, q - 5); // bytes spos'-5 .. spos'-2
                    const uint32_t t2 = ld_u32_unaligned(s_tile, q - 2); // bytes spos'-2 .. spos'+1
                    const uint32_t x3 = (t4 >> 16) & 0xffu, x4 = (t4 >> 8) & 0xffu, x5 = t4 & 0xffu;
                    const uint32_t hw = (x3 & 0x7fu) | ((x4 & 0x7fu) | is_alnum(x5) << 7) << 7;
                    words[hw] = (uint16_t)(t2 & 0xffffu);
                }
            }
        }
        after_literal = (it.symbol < 256) ? 1u : 0u;
        spos = new_spos;
        n_items++;
        __threadfence_block();
        __syncthreads();
    }

    __syncthreads();
    for (int i = tid; i < kNumCtx; i += kParseThreads) S->ctx_count[i] = s_cnt[i];
    for (int i = tid; i < kNumBuckets; i += kParseThreads) S->bhead[i] = s_bhead[i];
    if (tid == 0) {
        S->spos = spos;
        S->after_literal = after_literal;
        S->n_items = n_items;
    }
}


// LZEncoder::forward (lz.rs:82-87) in two grid-stride phases.
// Phase 1 — Bucket::forward (matcher.rs:82-87): rebase every position, saturating at 0.
__device__ __forceinline__ void enc_forward_rebase(ENode *nodes, uint32_t forward_len, int gid, int gsz) {
    for (int i = gid; i < kNumBuckets * kNodesStride; i += gsz) {
        if ((i & (kNodesStride - 1)) >= kRing) continue;
        const uint32_t pl = nodes[i].pl;
        const uint32_t p = node_pos(pl);
        nodes[i].pl = (p > forward_len ? p - forward_len : 0u) | (pl & 0xfe000000u);
    }
}
// Phase 2 — BucketMatcher::forward (matcher.rs:123-133): unlink heads / links whose target position became 0.
// The snapshot links next[1..] are cut at the first link whose target is dead: the reference would have cleared
// exactly that link in the predecessor's own slot (DESIGN.md stage A).
__device__ __forceinline__ void enc_forward_unlink(ENode *nodes, int16_t *heads, int gid, int gsz) {
    for (int i = gid; i < kNumBuckets * kHeadsStride; i += gsz) {
        const int b = i / kHeadsStride;
        if (i - b * kHeadsStride >= kHashSize) continue;
        const int h = heads[i];
        if (h != -1 && node_pos(nodes[b * kNodesStride + h].pl) == 0) heads[i] = -1;
    }
    for (int i = gid; i < kNumBuckets * kNodesStride; i += gsz) {
        if ((i & (kNodesStride - 1)) >= kRing) continue;
        const int b = i / kNodesStride;
        ENode *nd = &nodes[i];
        bool cut = false;
        for (int j = 0; j < kSnap; j++) {
            const int nx = nd->next[j];
            if (nx == -1) break;
            if (cut || node_pos(nodes[b * kNodesStride + nx].pl) == 0) {
                nd->next[j] = -1;
                cut = true;
            }
        }
    }
}

// ================================================================================================
// Stages B-D + framing + block loop
// ================================================================================================
constexpr int kEmitThreads = 512;
constexpr int kEmitWarps = kEmitThreads / 32;

// MSB-first bit writer into big-endian 32-bit words (coder.rs:159-217), single thread.
struct BitWriter {
    uint32_t *out;
    uint32_t wpos;
    uint64_t acc;
    int nbits;
    __device__ void init(uint32_t *o) {
        out = o;
        wpos = 0;
        acc = 0;
        nbits = 0;
    }
    __device__ void put(uint32_t bits, int len) {
        acc = (acc << len) | bits;
        nbits += len;
        if (nbits >= 32) {
            out[wpos++] = bswap32((uint32_t)(acc >> (nbits - 32)));
            nbits -= 32;
            acc &= (1ull << nbits) - 1ull;
        }
    }
    __device__ void varint(uint32_t v) { // coder.rs:27-38
        for (;;) {
            const uint32_t has_next = v > 1;
            put((v & 1u) | has_next << 1, 2);
            v >>= 1;
            if (!has_next) break;
        }
    }
    __device__ uint32_t total_bits() const { return wpos * 32u + (uint32_t)nbits; }
    __device__ void flush_partial() { // leaves the bit position unchanged; pads the current word with zeros
        if (nbits > 0) out[wpos] = bswap32((uint32_t)(acc << (32 - nbits)));
    }
};

// huffman.rs:27-111 for one table, executed by one thread on shared-memory scratch.
struct HuffScratch {
    uint32_t weight[2 * kSymStride];
    unsigned long long heap[kSymStride];
    uint16_t child1[2 * kSymStride];
    uint16_t child2[2 * kSymStride];
    uint8_t depth[2 * kSymStride];
};

__device__ void heap_push(unsigned long long *h, int &n, unsigned long long x) {
    int i = n++;
    while (i > 0) {
        const int p = (i - 1) >> 1;
        if (h[p] <= x) break;
        h[i] = h[p];
        i = p;
    }
    h[i] = x;
}
__device__ unsigned long long heap_pop(unsigned long long *h, int &n) {
    const unsigned long long top = h[0];
    const unsigned long long x = h[--n];
    int i = 0;
    for (;;) {
        int ch = 2 * i + 1;
        if (ch >= n) break;
        if (ch + 1 < n && h[ch + 1] < h[ch]) ch++;
        if (h[ch] >= x) break;
        h[i] = h[ch];
        i = ch;
    }
    if (n > 0) h[i] = x;
    return top;
}

// Keys (weight<<16 | index) are unique, so any correct min-heap pops in the reference's order.
__device__ void huffman_build(const uint32_t *sym_weights, int n, uint8_t *lens, HuffScratch &hs) {
    for (int i = 0; i < n; i++) hs.weight[i] = sym_weights[i];
    for (;;) {
        int hn = 0, nn = n;
        for (int i = 0; i < n; i++)
            if (sym_weights[i] > 0) heap_push(hs.heap, hn, ((unsigned long long)hs.weight[i] << 16) | (unsigned)i);
        if (hn <= 1) { // huffman.rs:64-71
            for (int i = 0; i < n; i++) lens[i] = 0;
            if (hn == 1) lens[(int)(hs.heap[0] & 0xffffu)] = 1;
            return;
        }
        while (hn > 1) { // huffman.rs:73-88
            const unsigned long long a = heap_pop(hs.heap, hn);
            const unsigned long long b = heap_pop(hs.heap, hn);
            const uint32_t wsum = (uint32_t)(a >> 16) + (uint32_t)(b >> 16);
            hs.weight[nn] = wsum;
            hs.child1[nn] = (uint16_t)(a & 0xffffu);
            hs.child2[nn] = (uint16_t)(b & 0xffffu);
            heap_push(hs.heap, hn, ((unsigned long long)wsum << 16) | (unsigned)nn);
            nn++;
        }
        for (int i = 0; i < nn; i++) hs.depth[i] = 0;
        for (int i = nn - 1; i >= n; i--) { // huffman.rs:91-96
            hs.depth[hs.child1[i]] = (uint8_t)(hs.depth[i] + 1);
            hs.depth[hs.child2[i]] = (uint8_t)(hs.depth[i] + 1);
        }
        int cur_max = 0;
        for (int i = 0; i < n; i++) cur_max = hs.depth[i] > cur_max ? hs.depth[i] : cur_max;
        if (cur_max > 15) { // huffman.rs:99-108
            const uint32_t shrink = 1u << (cur_max - 15);
            for (int i = 0; i < n; i++)
                if (hs.weight[i] > 0) {
                    const uint32_t wv = hs.weight[i] / shrink;
                    hs.weight[i] = wv > 1 ? wv : 1;
                }
            continue;
        }
        for (int i = 0; i < n; i++) lens[i] = hs.depth[i];
        return;
    }
}

// huffman.rs:119-141 canonical codes in (len, sym) order
__device__ void huffman_codes(const uint8_t *lens, int n, uint16_t *codes) {
    uint32_t bits = 0;
    int cur = 1;
    for (int len = 1; len <= 15; len++) {
        for (int s = 0; s < n; s++) {
            if (lens[s] != len) continue;
            if (len > cur) {
                bits <<= (len - cur);
                cur = len;
            }
            codes[s] = (uint16_t)bits;
            bits++;
        }
    }
}

// coder.rs:45-67
__device__ void write_huffman_table(BitWriter &bw, const uint8_t *lens, int n) {
    int maxl = 0;
    for (int i = 0; i < n; i++) maxl = lens[i] > maxl ? lens[i] : maxl;
    bw.varint((uint32_t)maxl);
    int last = -1;
    for (int s = 0; s < n; s++) {
        if (lens[s] > 0) {
            bw.varint(last < 0 ? (uint32_t)s + 1 : (uint32_t)(s - last));
            bw.varint((uint32_t)(maxl - lens[s]));
            last = s;
        }
    }
    bw.varint(0);
}

struct EmitShared {
    SymRank tab[kEmitWarps];          // per-warp rank table (symrank.rs)
    uint32_t hist1[2][kSymStride];    // lz.rs:272 huff_weights1
    uint32_t hist2[256];              // lz.rs:273 huff_weights2 (240 used)
    uint32_t base[kNumCtx];           // exclusive scan of ctx_count
    uint32_t cnt[kNumCtx];
    uint16_t order[kNumCtx];          // contexts by descending count
    uint32_t symcount[kSymStride];    // first chunk: pre-rank symbol counts (lz.rs:240-243)
    uint16_t vs[kSymStride];          // first chunk: symbol order (lz.rs:247-250)
    HuffScratch hs[3];
    uint8_t lens[3][kSymStride];
    uint16_t codes[3][kSymStride];
    uint32_t scan[kEmitThreads];
    uint32_t next_ctx;
    uint32_t num_counted;
    uint32_t hdr_bits;
    uint32_t total_bits;
};

__global__ void __launch_bounds__(kEmitThreads) orz_emit_kernel(EncStream *streams) {
    EncStream *S = &streams[blockIdx.x];
    if (S->status != 0) return;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    EmitShared &sh = *reinterpret_cast<EmitShared *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Item *items = S->items;
    uint32_t *sorted = S->sorted;
    const uint32_t n_items = S->n_items;
    const bool first_block = S->first_block != 0;

    // ---- init shared ----
    for (int i = tid; i < 2 * kSymStride; i += kEmitThreads) (&sh.hist1[0][0])[i] = 0;
    for (int i = tid; i < 256; i += kEmitThreads) sh.hist2[i] = 0;
    for (int i = tid; i < kSymStride; i += kEmitThreads) sh.symcount[i] = 0;
    for (int i = tid; i < kNumCtx; i += kEmitThreads) sh.cnt[i] = S->ctx_count[i];
    if (tid == 0) sh.next_ctx = 0;
    __syncthreads();

    // ---- first chunk of the stream: symbol order + rank table init (lz.rs:238-265) ----
    if (first_block) {
        for (uint32_t k = tid; k < n_items; k += kEmitThreads) atomicAdd(&sh.symcount[items[k].symbol], 1u);
        __syncthreads();
        if (tid < kNumSym) { // stable sort by Reverse(max(count,1)) via rank counting
            const uint32_t key = sh.symcount[tid] > 1 ? sh.symcount[tid] : 1;
            int rank = 0;
            for (int t = 0; t < kNumSym; t++) {
                const uint32_t kt = sh.symcount[t] > 1 ? sh.symcount[t] : 1;
                rank += (kt > key) || (kt == key && t < tid);
            }
            sh.vs[rank] = (uint16_t)tid;
        }
        if (tid == 0) sh.num_counted = 0;
        __syncthreads();
        if (tid < kNumSym && sh.symcount[tid] > 1) atomicAdd(&sh.num_counted, 1u);
        // init all 512 tables: value = vs, index = inverse, cnt = 0, sum = 1,000,000 (symrank.rs:22-36)
        for (int t = warp; t < kNumCtx; t += kEmitWarps) {
            SymRank *g = &S->symranks[t];
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
        __threadfence_block();
        __syncthreads();
    }

    // ---- group items by context: exclusive scan of the per-context counts, then scatter ----
    if (tid < kNumCtx) { // kEmitThreads == kNumCtx
        const uint32_t v = sh.cnt[tid];
        sh.scan[tid] = v;
    }
    __syncthreads();
    for (int off = 1; off < kNumCtx; off <<= 1) {
        uint32_t add = 0;
        if (tid >= off) add = sh.scan[tid - off];
        __syncthreads();
        sh.scan[tid] += add;
        __syncthreads();
    }
    sh.base[tid] = sh.scan[tid] - sh.cnt[tid];
    { // contexts by descending count (ties by index) for load balance
        const uint32_t key = sh.cnt[tid];
        int rank = 0;
        for (int t = 0; t < kNumCtx; t++) {
            const uint32_t kt = sh.cnt[t];
            rank += (kt > key) || (kt == key && t < tid);
        }
        sh.order[rank] = (uint16_t)tid;
    }
    __syncthreads();
    for (uint32_t k = tid; k < n_items; k += kEmitThreads) {
        const Item it = items[k];
        sorted[sh.base[it.ctx] + it.seq] = k;
    }
    __threadfence_block();
    __syncthreads();

    // ---- stage B: per-context rank coding + histograms (lz.rs:272-305, symrank.rs:38-97) ----
    {
        SymRank &tab = sh.tab[warp];
        for (;;) {
            uint32_t j = 0;
            if (lane == 0) j = atomicAdd(&sh.next_ctx, 1u);
            j = __shfl_sync(0xffffffffu, j, 0);
            if (j >= (uint32_t)kNumCtx) break;
            const int c = sh.order[j];
            const uint32_t cntc = sh.cnt[c];
            if (cntc == 0) break;
            SymRank *g = &S->symranks[c];
            {
                const uint32_t *src = reinterpret_cast<const uint32_t *>(g);
                uint32_t *dst = reinterpret_cast<uint32_t *>(&tab);
                for (int i = lane; i < (int)(sizeof(SymRank) / 4); i += 32) dst[i] = src[i];
            }
            __syncwarp();
            uint32_t ecnt = tab.cnt, esum = tab.sum;
            const uint32_t base = sh.base[c];
            for (uint32_t off = 0; off < cntc; off += 32) {
                const bool have = off + lane < cntc;
                uint32_t k = 0, sym = 0, unl = 0, flags = 0, eml = 0;
                if (have) {
                    k = sorted[base + off + lane];
                    const Item it = items[k];
                    sym = it.symbol;
                    unl = it.unlikely;
                    flags = it.flags;
                    eml = it.eml;
                }
                const int m = cntc - off < 32 ? (int)(cntc - off) : 32;
                uint32_t myrank = 0;
                for (int q = 0; q < m; q++) {
                    const uint32_t v = __shfl_sync(0xffffffffu, sym, q);
                    const uint32_t vu = __shfl_sync(0xffffffffu, unl, q);
                    const uint32_t i = tab.index[v];
                    const uint32_t iun = tab.index[vu];
                    // update (symrank.rs:61-97); every lane executes the same stores with the same values
                    if (ecnt > (uint32_t)kNumSym) {
                        ecnt = ecnt * 9 / 10;
                        esum = esum * 9 / 10;
                    }
                    ecnt += 1;
                    esum += i;
                    const uint32_t sub = ((i >> 4) + ((esum >> 4) / ecnt)) & 0xffffu; // u16 arithmetic
                    uint32_t next_i = i > sub ? i - sub : 0;
                    if (next_i < (i >> 1)) next_i = i >> 1;
                    const uint32_t d = i - next_i;
                    if (d == 1) {
                        const uint16_t nv1 = tab.value[next_i];
                        tab.index[v] = (uint16_t)next_i;
                        tab.value[i] = nv1;
                        tab.index[nv1] = (uint16_t)i;
                        tab.value[next_i] = (uint16_t)v;
                    } else if (d > 1) {
                        const uint32_t ni2 = next_i, ni1 = next_i + d / 2;
                        const uint16_t nv1 = tab.value[ni1];
                        const uint16_t nv2 = tab.value[ni2];
                        tab.value[i] = nv1;
                        tab.index[nv1] = (uint16_t)i;
                        tab.value[ni1] = nv2;
                        tab.index[nv2] = (uint16_t)ni1;
                        tab.value[ni2] = (uint16_t)v;
                        tab.index[v] = (uint16_t)ni2;
                    }
                    const uint32_t r = (i == iun) ? (uint32_t)(kNumSym - 1) : i - (i > iun); // symrank.rs:43-46
                    if (lane == q) myrank = r;
                    __syncwarp();
                }
                if (have) {
                    items[k].symbol = (uint16_t)myrank;
                    atomicAdd(&sh.hist1[flags & 1u][myrank], 1u);
                    if ((flags & 2u) && eml >= (uint32_t)(kLenidSize - 1)) atomicAdd(&sh.hist2[eml], 1u);
                }
            }
            __syncwarp();
            if (lane == 0) {
                tab.cnt = ecnt;
                tab.sum = esum;
            }
            __syncwarp();
            {
                uint32_t *dst = reinterpret_cast<uint32_t *>(g);
                const uint32_t *src = reinterpret_cast<const uint32_t *>(&tab);
                for (int i = lane; i < (int)(sizeof(SymRank) / 4); i += 32) dst[i] = src[i];
            }
            __syncwarp();
        }
    }
    __threadfence_block();
    __syncthreads();

    // ---- stage C: three length-limited Huffman tables (lz.rs:306-318), one thread each ----
    if (lane == 0 && warp < 3) {
        const uint32_t *wts = warp == 0 ? sh.hist1[0] : warp == 1 ? sh.hist1[1] : sh.hist2;
        const int n = warp < 2 ? kNumSym : kMatchMax;
        huffman_build(wts, n, sh.lens[warp], sh.hs[warp]);
        for (int i = 0; i < kSymStride; i++) sh.codes[warp][i] = 0;
        huffman_codes(sh.lens[warp], n, sh.codes[warp]);
    }
    __syncthreads();

    // ---- header (lz.rs:253-256,268-269,311-313), one thread ----
    uint32_t *pay = reinterpret_cast<uint32_t *>(S->payload);
    if (tid == 0) {
        BitWriter bw;
        bw.init(pay);
        if (first_block) {
            bw.varint(sh.num_counted);
            for (uint32_t k = 0; k < sh.num_counted; k++) bw.put(sh.vs[k], 9);
        }
        const uint32_t end_spos = S->spos < S->sbuf_len ? S->spos : S->sbuf_len;
        bw.varint(end_spos);
        bw.varint(n_items);
        write_huffman_table(bw, sh.lens[0], kNumSym);
        write_huffman_table(bw, sh.lens[1], kNumSym);
        write_huffman_table(bw, sh.lens[2], kMatchMax);
        bw.flush_partial();
        sh.hdr_bits = bw.total_bits();
    }
    __syncthreads();

    // ---- stage D: prefix scan of item bit lengths, big-endian bit-pack (lz.rs:320-344) ----
    const uint32_t per = (n_items + kEmitThreads - 1) / kEmitThreads;
    const uint32_t lo = min((uint32_t)tid * per, n_items), hi = min(lo + per, n_items);
    uint32_t mybits = 0;
    for (uint32_t k = lo; k < hi; k++) {
        const Item it = items[k];
        uint32_t b = sh.lens[it.flags & 1u][it.symbol];
        if (it.flags & 2u) {
            b += it.robitlen;
            if (it.eml >= kLenidSize - 1) b += sh.lens[2][it.eml];
        }
        mybits += b;
    }
    sh.scan[tid] = mybits;
    __syncthreads();
    for (int off = 1; off < kEmitThreads; off <<= 1) {
        uint32_t add = 0;
        if (tid >= off) add = sh.scan[tid - off];
        __syncthreads();
        sh.scan[tid] += add;
        __syncthreads();
    }
    const uint32_t hdr_bits = sh.hdr_bits;
    const uint32_t total_bits = hdr_bits + sh.scan[kEmitThreads - 1];
    const uint32_t start_bit = hdr_bits + sh.scan[tid] - mybits;
    const uint32_t total_words = (total_bits + 31) / 32;
    for (uint32_t wq = (hdr_bits + 31) / 32 + tid; wq < total_words; wq += kEmitThreads) pay[wq] = 0;
    __threadfence_block();
    __syncthreads();
    if (hi > lo && mybits > 0) {
        uint32_t wi = start_bit >> 5;
        int nb = (int)(start_bit & 31u); // virtual leading zero bits of the first word
        uint64_t acc = 0;
        bool first = nb != 0;
        for (uint32_t k = lo; k < hi; k++) {
            const Item it = items[k];
            const uint32_t al = it.flags & 1u;
            uint32_t fbits[3];
            int flen[3];
            fbits[0] = sh.codes[al][it.symbol];
            flen[0] = sh.lens[al][it.symbol];
            fbits[1] = it.robits;
            flen[1] = (it.flags & 2u) ? it.robitlen : 0;
            const bool long_len = (it.flags & 2u) && it.eml >= kLenidSize - 1;
            fbits[2] = long_len ? sh.codes[2][it.eml] : 0;
            flen[2] = long_len ? sh.lens[2][it.eml] : 0;
#pragma unroll
            for (int f = 0; f < 3; f++) {
                acc = (acc << flen[f]) | fbits[f];
                nb += flen[f];
                if (nb >= 32) {
                    const uint32_t word = bswap32((uint32_t)(acc >> (nb - 32)));
                    if (first) {
                        atomicOr(&pay[wi], word);
                        first = false;
                    } else {
                        pay[wi] = word;
                    }
                    wi++;
                    nb -= 32;
                    acc &= (1ull << nb) - 1ull;
                }
            }
        }
        if (nb > 0) atomicOr(&pay[wi], bswap32((uint32_t)(acc << (32 - nb))));
    }
    __threadfence_block();
    __syncthreads();

    const uint32_t tlen = total_words * 4; // coder.rs:75-82: padded to 32 bits
    if (tid == 0) {
        S->payload_len = tlen;
        S->first_block = 0;
        S->n_chunks += 1;
    }
    if (!S->self_driven) return;

    // ---- framing: ULEB128(t) then payload (lib.rs:79-80, ioutil.rs:79-88) ----
    uint64_t opos = S->out_pos;
    uint8_t hdr[5];
    int hl = 0;
    {
        uint32_t l = tlen;
        while (l >= 128) {
            hdr[hl++] = (uint8_t)(128 + (l & 127u));
            l >>= 7;
        }
        hdr[hl++] = (uint8_t)l;
    }
    const bool fits = opos + hl + tlen + 1 <= S->out_cap;
    if (fits) {
        uint8_t *dst = S->out + opos;
        if (tid < hl) dst[tid] = hdr[tid];
        dst += hl;
        const uint8_t *src = S->payload;
        for (uint32_t i = tid; i < tlen; i += kEmitThreads) dst[i] = src[i];
        opos += hl + tlen;
    }
    __syncthreads();

    // ---- block loop (lib.rs:72-86): next chunk, or slide + forward + next block, or EOF ----
    const uint32_t spos = S->spos, sbuf_len = S->sbuf_len;
    if (!fits) {
        if (tid == 0) S->status = -2; // ORZ_ENOSPC
        return;
    }
    if (spos < sbuf_len) {
        if (tid == 0) S->out_pos = opos;
        return;
    }
    const uint64_t in_pos = S->in_pos, in_len = S->in_len;
    if (in_pos >= in_len) { // lib.rs:89: empty chunk marks EOF
        if (tid == 0) {
            S->out[opos] = 0;
            S->out_pos = opos + 1;
            S->status = 1;
        }
        return;
    }
    // slide: copy_within(len-PREMATCH.., 0) (lib.rs:83); the window is the full 33,554,431-byte slice here
    uint8_t *win = S->win;
    for (uint32_t i = tid; i < (kPrematch + 15) / 16; i += kEmitThreads) {
        // source and destination are 16,777,216 apart and 16 B aligned; ranges do not overlap
        const uint4 v = *reinterpret_cast<const uint4 *>(win + kBlockNew + 16u * i);
        if (16u * i + 16u <= kPrematch) {
            *reinterpret_cast<uint4 *>(win + 16u * i) = v;
        } else {
            const uint8_t *vb = reinterpret_cast<const uint8_t *>(&v);
            for (uint32_t b = 0; 16u * i + b < kPrematch; b++) win[16u * i + b] = vb[b];
        }
    }
    // forward (lz.rs:82-87, matcher.rs:82-87,123-133)
    ENode *nodes = S->nodes;
    int16_t *heads = S->heads;
    enc_forward_rebase(nodes, kBlockNew, tid, kEmitThreads);
    __threadfence_block();
    __syncthreads();
    enc_forward_unlink(nodes, heads, tid, kEmitThreads);
    __syncthreads();
    // read the next block into [PREMATCH, PREMATCH+rd) (lib.rs:72)
    const uint64_t rem = in_len - in_pos;
    const uint32_t rd = rem < kBlockNew ? (uint32_t)rem : kBlockNew;
    const uint8_t *src = S->in + in_pos;
    for (uint32_t i = tid; i < rd; i += kEmitThreads) win[kPrematch + i] = src[i];
    if (tid == 0) {
        S->out_pos = opos;
        S->in_pos = in_pos + rd;
        S->spos = kPrematch;
        S->sbuf_len = kPrematch + rd;
    }
}

// First block of every stream (lib.rs:67-75): window is zero except the new data.
__global__ void __launch_bounds__(256) orz_enc_begin_kernel(EncStream *streams) {
    EncStream *S = &streams[blockIdx.x];
    const int tid = threadIdx.x;
    const uint64_t in_len = S->in_len;
    if (in_len == 0) { // lib.rs:72-74,89: nothing read, only the EOF marker
        if (tid == 0) {
            if (S->out_cap >= 1) {
                S->out[0] = 0;
                S->out_pos = 1;
                S->status = 1;
            } else {
                S->status = -2;
            }
        }
        return;
    }
    const uint32_t rd = in_len < kBlockNew ? (uint32_t)in_len : kBlockNew;
    uint8_t *win = S->win;
    const uint8_t *src = S->in;
    for (uint32_t i = tid; i < rd; i += 256) win[kPrematch + i] = src[i];
    if (tid == 0) {
        S->in_pos = rd;
        S->spos = kPrematch;
        S->sbuf_len = kPrematch + rd;
        S->status = 0;
    }
}

// LZEncoder::forward for the chunk-level API (lz.rs:82-87)
__global__ void __launch_bounds__(256) orz_enc_forward_kernel(EncStream *streams, uint32_t forward_len, int phase) {
    EncStream *S = &streams[0];
    const int gid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    if (phase == 0) enc_forward_rebase(S->nodes, forward_len, gid, gsz);
    else enc_forward_unlink(S->nodes, S->heads, gid, gsz);
}

size_t emit_smem_bytes() { return sizeof(EmitShared); }

void launch_parse(EncStream *d_streams, int n, cudaStream_t st) { orz_parse_kernel<<<n, kParseThreads, 0, st>>>(d_streams); }
void launch_emit(EncStream *d_streams, int n, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(orz_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EmitShared));
        attr_set = true;
    }
    orz_emit_kernel<<<n, kEmitThreads, sizeof(EmitShared), st>>>(d_streams);
}
void launch_enc_begin(EncStream *d_streams, int n, cudaStream_t st) { orz_enc_begin_kernel<<<n, 256, 0, st>>>(d_streams); }
void launch_enc_forward(EncStream *d_stream, uint32_t forward_len, cudaStream_t st) {
    orz_enc_forward_kernel<<<148, 256, 0, st>>>(d_stream, forward_len, 0);
    orz_enc_forward_kernel<<<148, 256, 0, st>>>(d_stream, forward_len, 1);
}

} // namespace orz
