// encode_kernels.cu — sm_100a kernels of the ROLZ encoder.
//
//   orz_parse_kernel   stage A  (lz.rs:131-235): one thread block owns one stream.  The ROLZ parse is a serial
//                      state machine; builder warps speculatively gather and resolve the match candidates of the
//                      next positions (ring snapshots -> one parallel gather per chain, ballot-resolved semantics of
//                      matcher.rs:135-228) under per-bucket seqlocks, one executor thread takes the decisions.
//   orz_emit_kernel    stages B-D (lz.rs:238-344): one 512-thread block per stream — per-context rank
//                      coding (one warp per context, table in shared memory), shared-memory histograms,
//                      length-limited canonical Huffman build, block-wide prefix scan of bit lengths and
//                      big-endian bit-pack, framing (lib.rs:79-80) and the block loop (lib.rs:72-86).
//
// Integer/byte work, latency bound: no tensor cores.  All reference citations are file:line into
// richox/orz @29414c6.
#include "orz_dev.cuh"

#include <cstdlib>
#include <cstring>

namespace orz {

// ================================================================================================
// Stage A  (lz.rs:131-235)
//
// The parse is one serial chain per stream: which positions get inserted into the rings depends on
// every earlier decision.  What CAN run ahead of the chain is the expensive, read-only part of a step —
// hashing, the chain gather, the LCPs and the resolution of matcher.rs:135-228 — as long as the bucket it
// read is unchanged when the chain arrives there.  So one block per stream is split into
//   * builder warps: builder b speculatively builds a "probe" for the next position q >= spos with
//     q % NB == b (all candidate records gathered in parallel through the ring snapshots, LCPs against the
//     window, find_match resolved, lazy maxima), under a per-bucket seqlock, and publishes a summary in
//     shared memory;
//   * one executor warp (alone on its SM sub-partition): consumes the probes of spos, spos+1, spos+2,
//     takes the serial decision, writes the ring slot / head / item, bumps the bucket version and moves on.
// A probe whose bucket was written after it was built is rebuilt; a probe of a skipped position is dropped.
// ================================================================================================
constexpr unsigned kFull = 0xffffffffu;
constexpr uint32_t kNone = 0xffffffffu;

// What a builder publishes about position q (shared memory).  Everything the serial decision needs is
// precomputed here, so the executor only selects, stamps and stores.
struct __align__(16) ProbePub {
    // vector 0 — polled by the executor
    uint32_t q;        // position described, kNone while the slot is being rewritten (written last)
    uint32_t ver;      // bver[c] the probe was built under (the executor may re-stamp it, see restamp_probe)
    uint32_t cmx;      // c | min(max_d1,255) << 8 | min(max_d2,255) << 16 | single_round << 24
    uint32_t hc;       // hash2(q-1) | bytes(q,q+1) << 16
    // vector 1 — the match outcome (valid while the bucket is unchanged)
    uint32_t fm;       // fm_len | fm_idx << 8 | bump << 20      (fm_len 0 = no match)
    uint32_t sym;      // match symbol | robits << 16
    uint32_t aux;      // robitlen | eml << 8 | lazy_len1 << 16
    uint32_t bump_pl;  // new pos|lenmin of the matched slot if bump
    // vector 2
    uint32_t e0;       // hash_dword(q) % 4627
    uint32_t w_lit;    // words update if the step emits a literal: index | value << 16
    uint32_t w_m;      // words update if the step emits the match
    uint32_t fm_plexp; // raw lenmin << 8 | raw expected of the matched slot (to redo `aux` after a re-stamp)
    uint32_t rec[16];  // image of the ring slot this step writes: pl=q, expected (patched), next[0]=h0, next[1..]=old head's snapshot
    uint16_t cand[32]; // slots of the chain nodes that were read (first round), 0xffff = none
};

// Reads the head record of a chain: pl, expected, and (lane L in 1..kSnap) the snapshot link next[L-1].
__device__ __forceinline__ void enode_head(const ENode *bn, int slot, uint32_t &pl, uint32_t &exp, int &my_next, int lane) {
    const uint32_t *r = reinterpret_cast<const uint32_t *>(bn + slot);
    pl = __shfl_sync(0xffffffffu, r[0], 0);
    exp = __shfl_sync(0xffffffffu, r[1], 0) & 0xffu;
    my_next = -1;
    if (lane >= 1 && lane <= kSnap) {
        const uint32_t w = r[(lane + 2) >> 1]; // int16 index lane+2 of the record
        my_next = (int)(int16_t)(((lane + 2) & 1) ? (w >> 16) : (w & 0xffffu));
    }
}

// Longest common prefix of window[npos..] and window[qpos..], capped at 240
// (mem.rs:41-51: 16-byte steps, same result as a byte-wise LCP capped at max_len).
__device__ __forceinline__ int lcp_win_win(const uint8_t *win, uint32_t npos, uint32_t qpos) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(win + (npos & ~3u));
    const uint32_t sh = (npos & 3u) * 8u;
    const uint32_t *t = reinterpret_cast<const uint32_t *>(win + (qpos & ~3u));
    const uint32_t tsh = (qpos & 3u) * 8u;
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

__device__ __forceinline__ void sts_volatile_v4(volatile void *p, uint4 v) {
#ifdef ORZ_SIMT_EMU
    *reinterpret_cast<uint4 *>(const_cast<void *>(p)) = v;
    return;
#endif
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(const_cast<void *>(p));
    asm volatile("st.volatile.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// lane 0's value for the whole warp (also reconverges the warp)
__device__ __forceinline__ uint32_t bcast0(uint32_t v) { return __shfl_sync(kFull, v, 0); }

__device__ __forceinline__ int warp_max(int v) {
    for (int o = 16; o; o >>= 1) {
        const int other = __shfl_xor_sync(kFull, v, o);
        v = other > v ? other : v;
    }
    return v;
}

struct ProbeResult {
    int fm_len, fm_idx;
    uint32_t fm_pl, fm_exp;
    int max_d1, max_d2;
};

// matcher.rs:135-192 (find_match at q, depth0) and matcher.rs:194-228 restated as "the longest LCP among the
// first depth1 / depth2 chain nodes" (has_lazy_match(min_len) <=> that maximum >= min_len), from ONE walk of the
// chain that starts at h0.  f_next: lane L (1..kSnap) returns the head node's snapshot link next[L-1].
// Returns false if the walk was abandoned because the parse already moved past q.
__device__ __forceinline__ bool probe_chain(const uint8_t *win, const ENode *bn, int h0, uint32_t q, uint32_t buf_len, int depth0,
                                            int depth1, int depth2, int lane, const volatile uint32_t *s_spos, ProbeResult &R,
                                            int &f_next, int &cand_slot, bool &single_round) {
    R.fm_len = 0;
    R.fm_idx = 0;
    R.fm_pl = 0;
    R.fm_exp = 0;
    R.max_d1 = -1;
    R.max_d2 = -1;
    f_next = -1;
    cand_slot = -1;
    single_round = true;
    if (h0 < 0) return true;
    uint32_t x_pl, x_exp;
    int x_idx = h0, my_next;
    enode_head(bn, h0, x_pl, x_exp, my_next, lane);
    f_next = my_next;
    int dmax = depth0 > depth1 ? depth0 : depth1;
    dmax = dmax > depth2 ? dmax : depth2;
    if (dmax <= 0) return true;

    int max_len = kMatchMin - 1;
    uint32_t exp_best = kMatchMin, pl_best = (uint32_t)kMatchMin << 25;
    int idx_best = 0;
    int seen = 0;          // chain nodes consumed by earlier rounds
    bool fm_done = depth0 <= 0;
    bool first_round = true;

    for (;;) {
        if (bcast0(*s_spos) > q) return false;
        const int lo = first_round ? 0 : 1; // the head node is a candidate only in the first round
        const int room = dmax - seen - (first_round ? 1 : 0);
        const int cap = room < kSnap ? room : kSnap;
        uint32_t my_pl, my_exp;
        int my_idx;
        const int nvl = gather_round(bn, lane, cap, x_pl, x_exp, x_idx, my_next, my_pl, my_exp, my_idx);
        const int gi = seen + lane - lo;                 // index of this lane's node in chain order
        const bool have = lane >= lo && lane < nvl;
        if (first_round) cand_slot = (lane < nvl || (lane == nvl && my_idx >= 0)) ? my_idx : -1; // include the slot the walk stopped at
        else single_round = false;
        int lcp = -1;
        if (have) lcp = lcp_win_win(win, node_pos(my_pl), q);

        R.max_d1 = max(R.max_d1, warp_max((have && gi < depth1) ? lcp : -1));
        R.max_d2 = max(R.max_d2, warp_max((have && gi < depth2) ? lcp : -1));

        if (!fm_done) { // resolve the sequential semantics of matcher.rs:149-181 over this round's candidates
            const bool valid = have && gi < depth0;
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
                    if ((brk >> lane) & 1u)
                        eq = ld_u32_unaligned(win, node_pos(my_pl) + max_len - 3) == ld_u32_unaligned(win, q + max_len - 3);
                    if (__ballot_sync(kFull, eq)) {
                        fm_done = true;
                        break;
                    }
                }
                if (u >= 32) break;
                max_len = __shfl_sync(kFull, lcp, u);
                exp_best = __shfl_sync(kFull, my_exp, u);
                pl_best = __shfl_sync(kFull, my_pl, u);
                idx_best = __shfl_sync(kFull, my_idx, u);
                if (max_len == kMatchMax || (exp_best > 0 && max_len > (int)exp_best)) { // matcher.rs:162-167
                    fm_done = true;
                    break;
                }
                cur = u;
            }
        }
        seen += nvl - lo;
        if (seen >= depth0) fm_done = true;
        const bool lazy_done = (seen >= depth1 || R.max_d1 >= kMatchMax / 2 + 2) && (seen >= depth2 || R.max_d2 >= kMatchMax / 2 + 2);
        if (nvl <= cap || (fm_done && lazy_done)) break; // nvl <= cap: the chain ended inside this round
        // continue from the last node of this round (lane kSnap): it becomes the next round's head
        x_pl = __shfl_sync(kFull, my_pl, kSnap);
        x_exp = __shfl_sync(kFull, my_exp, kSnap);
        x_idx = __shfl_sync(kFull, my_idx, kSnap);
        uint32_t d0, d1;
        enode_head(bn, x_idx, d0, d1, my_next, lane);
        first_round = false;
    }
    if (max_len >= kMatchMin && q + max_len < buf_len) { // matcher.rs:183
        R.fm_len = max_len;
        R.fm_idx = idx_best;
        R.fm_pl = pl_best;
        R.fm_exp = exp_best;
    }
    return true;
}

// lz.rs:148,151-153,173-180 for a found match, given the ring head it will be coded against.
// Returns sym = symbol | robits << 16 and aux = robitlen | eml << 8 | lazy_len1 << 16.
__device__ __forceinline__ void match_fields(uint32_t head, uint32_t fm_idx, uint32_t fm_len, uint32_t raw_lenmin, uint32_t raw_exp,
                                             uint32_t &sym, uint32_t &aux) {
    const uint32_t ro = (head + kRing - fm_idx) % kRing; // matcher.rs:185
    uint32_t roid, robitlen, robits;
    roid_encode(ro, roid, robitlen, robits);             // lz.rs:148
    const uint32_t mexp = raw_exp > (uint32_t)kMatchMin ? raw_exp : (uint32_t)kMatchMin;
    const uint32_t mmin = raw_lenmin > (uint32_t)kMatchMin ? raw_lenmin : (uint32_t)kMatchMin;
    uint32_t eml;
    if (fm_len > mexp) eml = fm_len - mmin;
    else if (fm_len < mexp) eml = fm_len - mmin + 1;
    else eml = 0;
    eml &= 0xffu; // `as u8`
    const uint32_t lenid = eml < (uint32_t)(kLenidSize - 1) ? eml : (uint32_t)(kLenidSize - 1);
    const uint32_t lazy_len1 = fm_len + 1 + (robitlen < 8);
    sym = (256 + roid * kLenidSize + lenid) | robits << 16;
    aux = robitlen | eml << 8 | lazy_len1 << 16;
}

constexpr int kMaxBuilders = 16;

// one 16-byte volatile read of shared memory (the executor's poll)
__device__ __forceinline__ uint4 lds_volatile_v4(const volatile void *p) {
    uint4 v;
#ifdef ORZ_SIMT_EMU
    ORZ_POLL();
    v = *reinterpret_cast<const uint4 *>(const_cast<const void *>(p));
    return v;
#endif
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(const_cast<void *>(p));
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}

__device__ __forceinline__ uint2 lds_volatile_v2(const volatile void *p) {
    uint2 v;
#ifdef ORZ_SIMT_EMU
    ORZ_POLL();
    v = *reinterpret_cast<const uint2 *>(const_cast<const void *>(p));
    return v;
#endif
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(const_cast<void *>(p));
    asm volatile("ld.volatile.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
    return v;
}

struct ParseShared {
    uint16_t words[32768];         // lz.rs:52 — only the executor touches it
    uint32_t cnt[kNumCtx];
    uint32_t bver[kNumBuckets];    // per-bucket seqlock: odd while the executor is writing the bucket
    uint16_t bhead[kNumBuckets];
    uint32_t spos;                 // published parse position
    uint32_t done;
    ProbePub pub[2 * kMaxBuilders]; // slot of position q: q % (2 * NB); builder b owns the positions q % NB == b
};

// The executor (one thread) just inserted (slot new_head, hash entry e0, optional lenmin bump of slot bump_idx)
// into bucket c and holds its seqlock.  If the published probe P lives in the same bucket but read none of what
// changed, it is still exact: re-stamp it with the bucket's next version instead of letting its builder redo the walk.
__device__ __forceinline__ void restamp_probe(volatile ProbePub *P, uint32_t want_q, uint32_t c, uint32_t old_ver, uint32_t e0,
                                              uint32_t new_head, int bump_idx) {
    const uint4 v = lds_volatile_v4(P);
    if (v.x != want_q || v.y != old_ver || (v.z & 0xffu) != c || !((v.z >> 24) & 1u) || P->e0 == e0) return;
    const uint32_t fm = P->fm;
    const uint32_t fm_len = fm & 0xffu, fm_idx = (fm >> 8) & 0xfffu;
    if (fm_len > 0 && bump_idx >= 0 && fm_idx == (uint32_t)bump_idx) return;
    const uint32_t pat = new_head | new_head << 16;
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint4 cw = lds_volatile_v4(reinterpret_cast<volatile uint32_t *>(const_cast<uint16_t *>(P->cand)) + 4 * k);
        any |= __vcmpeq2(cw.x, pat) | __vcmpeq2(cw.y, pat) | __vcmpeq2(cw.z, pat) | __vcmpeq2(cw.w, pat);
    }
    if (any) return; // the probe read the slot that was just overwritten
    if (fm_len > 0) { // the ring head moved by one: the reduced offset of P's match changes with it
        const uint32_t pe = P->fm_plexp;
        uint32_t sym, aux;
        match_fields(new_head, fm_idx, fm_len, pe >> 8, pe & 0xffu, sym, aux);
        P->sym = sym;
        P->aux = aux;
        const uint32_t nh2 = new_head + 1 == kRing ? 0 : new_head + 1;
        const bool bump = (pe >> 8) <= fm_len && fm_idx != nh2;
        P->fm = (fm & 0xfffffu) | (bump ? 1u << 20 : 0u);
    }
    P->ver = old_ver + 2;
}

// NW warps per block.  Warp 0 is the executor; warps with (warp & 3) == 0 share its SM sub-partition and stay
// parked so that it issues alone; the other NB warps are builders (NW = 22 -> 16 builders, NW = 11 -> 8).
template <int NW, int NB, bool STATS> // STATS: compile the counters / clock64 instrumentation in (ORZ_B200_STATS); off in production
__global__ void __launch_bounds__(NW * 32) orz_parse_kernel(EncStream *streams) {
    static_assert(NB == NW - 1 - (NW - 1) / 4 && NB <= kMaxBuilders, "builder count");
    EncStream *S = &streams[blockIdx.x];
    if (S->status != 0) return;
    ORZ_DYN_SHARED(smem_raw);
    ParseShared &sh = *reinterpret_cast<ParseShared *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NT = NW * 32;
    constexpr int NS = 2 * NB; // probe slots = lookahead in positions

    for (int i = tid; i < kNumCtx; i += NT) sh.cnt[i] = 0;
    for (int i = tid; i < kNumBuckets; i += NT) {
        sh.bhead[i] = (uint16_t)S->bhead[i];
        sh.bver[i] = 0;
    }
    for (int i = tid; i < 32768 / 8; i += NT) reinterpret_cast<uint4 *>(sh.words)[i] = reinterpret_cast<const uint4 *>(S->words)[i];
    if (tid < 2 * NB) sh.pub[tid].q = kNone;
    const uint8_t *win = S->win;
    ENode *nodes = S->nodes;
    int16_t *heads = S->heads;
    const uint32_t sbuf_len = S->sbuf_len;
    const int depth0 = (int)S->depth0, depth1 = (int)S->depth1, depth2 = (int)S->depth2;
    const uint32_t dbg_flags = S->dbg_flags;
    if (tid == 0) {
        sh.spos = S->spos;
        sh.done = 0;
    }
    __syncthreads();
    volatile uint32_t *v_spos = &sh.spos;
    volatile uint32_t *v_done = &sh.done;
    volatile uint32_t *v_bver = sh.bver;
    volatile uint16_t *v_bhead = sh.bhead;

    if (warp == 0) {
        // ------------------------------------------------------------------ executor (ONE thread: every poll, branch and
        // store of the serial chain is taken by lane 0, so no decision can differ between lanes)
        if (lane == 0) {
            Item *items = S->items;
            const uint32_t items_cap = S->items_cap < kChunkItems ? S->items_cap : kChunkItems;
            uint32_t spos = S->spos;
            uint32_t after_literal = S->after_literal;
            uint32_t n_items = 0;
            unsigned long long w0 = 0, w1 = 0, w2 = 0;
            unsigned cz_building = 0, cz_other = 0, cz_stale = 0;
            long long cyc_own = 0, cyc_lazy = 0;
            const long long t_start = STATS ? clock64() : 0;
            // watchdog: a probe that never arrives is a protocol bug; fail the stream instead of hanging the GPU
            const unsigned wd_limit = 30000000u; // spin iterations of ONE wait (seconds); normal waits are < 100
            int wd_code = 0;
            while (spos < sbuf_len && n_items < items_cap) { // lz.rs:131
                volatile ProbePub *P = &sh.pub[spos % NS];
                uint32_t c, ver, hc;
                unsigned wl = 0;
                const long long t_a = STATS ? clock64() : 0;
                for (;;) { // wait for a probe of spos built under the current version of its bucket
                    if (STATS) w0++;
                    if (++wl > wd_limit) {
                        wd_code = 1;
                        break;
                    }
                    const uint4 v = lds_volatile_v4(P);
                    if (v.x != spos) {
                        if (STATS && wl == 1) (v.x == kNone ? cz_building : cz_other)++;
                        continue;
                    }
                    ver = v.y;
                    hc = v.w;
                    c = v.z & 0xffu;
                    if (v_bver[c] == ver) break;
                    if (STATS && wl == 1) cz_stale++;
                }
                if (wd_code) break;
                const long long t_b = STATS ? clock64() : 0;
                if (STATS) cyc_own += t_b - t_a;
                const uint4 v1 = lds_volatile_v4(&P->fm);
                const uint4 v2 = lds_volatile_v4(&P->e0);
                uint4 r0 = lds_volatile_v4(&P->rec[0]);
                const uint4 r1 = lds_volatile_v4(&P->rec[4]), r2 = lds_volatile_v4(&P->rec[8]), r3 = lds_volatile_v4(&P->rec[12]);
                volatile ProbePub *P1 = &sh.pub[(spos + 1) % NS];
                volatile ProbePub *P2 = &sh.pub[(spos + 2) % NS];
                uint4 a1 = lds_volatile_v4(P1), a2 = lds_volatile_v4(P2); // lazy probes: read now, validated only if needed
                const uint32_t h2 = hc & 0x7fffu, cur2 = hc >> 16;
                const uint32_t fm_len = v1.x & 0xffu, fm_idx = (v1.x >> 8) & 0xfffu;
                const bool bump = (v1.x >> 20) & 1u;
                const uint32_t e0 = v2.x;

                const uint32_t lwe = sh.words[h2];                       // lz.rs:132
                const bool last_word_matched = cur2 == lwe;              // lz.rs:133
                const uint32_t ctx = c | after_literal << 8;             // lz.rs:134-135
                int lazy_id = 0;
                if (fm_len > 0 && fm_len < kMatchMax / 2) {              // lz.rs:151-171
                    const uint32_t lazy_len1 = (v1.z >> 16) & 0xffu;
                    const uint32_t lazy_len2 = lazy_len1 - (uint32_t)last_word_matched;
                    wl = 0;
                    for (;;) {
                        if (STATS) w1++;
                        if (++wl > wd_limit) {
                            wd_code = 2;
                            break;
                        }
                        if (a1.x == spos + 1 && v_bver[a1.z & 0xffu] == a1.y) break;
                        a1 = lds_volatile_v4(P1);
                    }
                    if (((a1.z >> 8) & 0xffu) >= lazy_len1) {
                        lazy_id = 1;
                    } else {
                        wl = 0;
                        for (;;) {
                            if (STATS) w2++;
                            if (++wl > wd_limit) {
                                wd_code = 3;
                                break;
                            }
                            if (a2.x == spos + 2 && v_bver[a2.z & 0xffu] == a2.y) break;
                            a2 = lds_volatile_v4(P2);
                        }
                        if (((a2.z >> 16) & 0xffu) >= lazy_len2) lazy_id = 2;
                    }
                }
                if (wd_code) break;
                if (STATS) cyc_lazy += clock64() - t_b;

                const bool emit_match = fm_len > 0 && lazy_id == 0;
                const uint32_t mlen = emit_match ? fm_len : 0u;
                const uint32_t seq = sh.cnt[ctx];
                uint32_t new_spos, wupd;
                Item it;
                it.ctx = (uint16_t)ctx;
                it.unlikely = (uint8_t)(lwe & 0xffu);                // lz.rs:136
                it.pad = 0;
                it.seq = seq;
                if (emit_match) { // lz.rs:173-209
                    it.symbol = (uint16_t)(v1.y & 0xffffu);
                    it.robits = (uint16_t)(v1.y >> 16);
                    it.robitlen = (uint8_t)(v1.z & 0xffu);
                    it.eml = (uint8_t)((v1.z >> 8) & 0xffu);
                    it.flags = (uint8_t)(after_literal | 2u);
                    new_spos = spos + mlen;
                    wupd = v2.z;
                } else {
                    it.robits = 0;
                    it.robitlen = 0;
                    it.eml = 0;
                    it.flags = (uint8_t)after_literal;
                    wupd = v2.y;
                    if (spos + 1 < sbuf_len && lazy_id != 1 && last_word_matched) { // lz.rs:218-226
                        it.symbol = (uint16_t)kWordSym;
                        new_spos = spos + 2;
                    } else { // lz.rs:227-236
                        it.symbol = (uint16_t)(cur2 & 0xffu);
                        new_spos = spos + 1;
                    }
                }
                *v_spos = new_spos; // builders may retarget now: everything of P is in registers
                const int bump_idx = (emit_match && bump) ? (int)fm_idx : -1;

                // Bucket::update (matcher.rs:62-80) + BucketMatcher::update (matcher.rs:115-121) under the bucket's seqlock
                ENode *bn = nodes + c * kNodesStride;
                const uint32_t head = sh.bhead[c];
                const uint32_t new_head = head + 1 == kRing ? 0 : head + 1;
                v_bver[c] = ver + 1;
                __threadfence_block();
                if (emit_match && bump) bn[fm_idx].pl = v1.w;
                r0.y = (r0.y & 0xffffff00u) | mlen; // expected = match_len (matcher.rs:75)
                uint4 *dst = reinterpret_cast<uint4 *>(bn + new_head);
                dst[0] = r0;
                dst[1] = r1;
                dst[2] = r2;
                dst[3] = r3;
                heads[c * kHeadsStride + e0] = (int16_t)new_head;
                v_bhead[c] = (uint16_t)new_head;
                // neighbours in the same bucket that did not read anything this step changed stay valid
                if (!(dbg_flags & 1u)) {
                    restamp_probe(&sh.pub[new_spos % NS], new_spos, c, ver, e0, new_head, bump_idx);
                    restamp_probe(&sh.pub[(new_spos + 1) % NS], new_spos + 1, c, ver, e0, new_head, bump_idx);
                    restamp_probe(&sh.pub[(new_spos + 2) % NS], new_spos + 2, c, ver, e0, new_head, bump_idx);
                }
                __threadfence_block();
                v_bver[c] = ver + 2;
                sh.cnt[ctx] = seq + 1;
                items[n_items] = it;
                if (it.symbol != kWordSym) sh.words[wupd & 0xffffu] = (uint16_t)(wupd >> 16); // lz.rs:208,233
                after_literal = (it.symbol < 256) ? 1u : 0u;
                spos = new_spos;
                n_items++;
            }
            *v_done = 1;
            if (wd_code) {
                S->status = -7;
                S->dbg[6] = (unsigned long long)wd_code;
                S->dbg[7] = spos;
            }
            S->spos = spos;
            S->after_literal = after_literal;
            S->n_items = n_items;
            if (STATS) S->dbg[0] += n_items;
            if (STATS && (dbg_flags & 8u)) { // report cycles instead of spin counts
                w0 = (unsigned long long)cyc_own;
                w1 = (unsigned long long)cyc_lazy;
                w2 = (unsigned long long)(clock64() - t_start);
            }
            if (STATS && !(dbg_flags & 32u)) {
                S->dbg[1] += w0;
                S->dbg[2] += w1;
                S->dbg[3] += w2;
            }
            if (STATS && !wd_code) {
                S->dbg[6] += ((unsigned long long)cz_building << 32) | cz_other;
                S->dbg[7] += cz_stale;
            }
        }
    } else if ((warp & 3) != 0) {
        // ------------------------------------------------------------------ builders
        const int b = warp - 1 - (warp >> 2);
        // this builder's two published probes (positions b mod NB, NB apart): [0] lives in slot q % NS < NB ... by position
        uint32_t pq[2] = {kNone, kNone}, pc[2] = {0, 0}, pver[2] = {0, 0};
        unsigned long long nbuild = 0, nabort = 0;
        unsigned long long bcyc = 0, nvfail = 0, lead[4] = {0, 0, 0, 0};
        for (;;) {
            // Control values are read by every lane but only lane 0's copy is used (bcast0): lanes that are not
            // perfectly converged could otherwise see different values and part ways before a full-mask shuffle.
            if (bcast0(*v_done)) break;
            const uint32_t sp = bcast0(*v_spos);
            const uint32_t q1 = sp + ((uint32_t)b + NB - sp % NB) % NB; // next position >= sp owned by this builder
            // pick the nearer of its two positions that has no current probe
            uint32_t q = kNone;
            int k = 0;
#pragma unroll
            for (int t = 0; t < 2; t++) {
                const uint32_t qt = q1 + t * NB;
                const int kt = (int)((qt / NB) & 1u);
                bool ok = qt >= sbuf_len;
                if (!ok && pq[kt] == qt) {
                    const uint32_t bv = bcast0(v_bver[pc[kt]]);
                    if (bv == pver[kt]) ok = true;                 // published and still current
                    else if (bv & 1u) ok = true;                   // the executor is writing the bucket (may re-stamp): look again later
                    else {
                        volatile ProbePub *vp = &sh.pub[qt % NS];
                        const uint32_t pv = bcast0(vp->ver), pqq = bcast0(vp->q);
                        if (pv == bv && pqq == qt) {               // re-stamped by the executor: still exact
                            pver[kt] = bv;
                            ok = true;
                        }
                    }
                }
                if (!ok && q == kNone) {
                    q = qt;
                    k = kt;
                }
            }
            if (q == kNone) {
                if (dbg_flags & 64u) __nanosleep(100); // ORZ_B200_DBG=64: idle builders sleep instead of spinning (to be measured with many streams per SM)
                continue;
            }
            ProbePub &pub = sh.pub[q % NS];
            volatile ProbePub *vpub = &pub;
            uint32_t &my_q = pq[k], &my_c = pc[k], &my_ver = pver[k];
            // ---- build the probe of q ----
            my_q = kNone;
            if (STATS) nbuild++;
            const long long tb0 = STATS ? clock64() : 0;
            if (lane == 0) vpub->q = kNone;
            __threadfence_block();
            const uint32_t prev4 = ld_u32_unaligned(win, q - 4); // bytes q-4 .. q-1
            const uint32_t cur4 = ld_u32_unaligned(win, q);      // bytes q .. q+3
            const uint32_t b_m1 = prev4 >> 24, b_m2 = (prev4 >> 16) & 0xffu, b_m3 = (prev4 >> 8) & 0xffu, b_m4 = prev4 & 0xffu;
            const uint32_t c = (b_m1 & 0x7fu) | is_alnum(b_m2) << 7;                              // hash1(q-1)  lz.rs:481-486
            const uint32_t h2 = (b_m1 & 0x7fu) | ((b_m2 & 0x7fu) | is_alnum(b_m3) << 7) << 7;     // hash2(q-1)  lz.rs:488-492
            const uint32_t e0 = hash_dword(cur4) % kHashSize;
            const uint32_t v0 = bcast0(v_bver[c]);
            if (v0 & 1u) continue; // the executor is writing this bucket right now
            __threadfence_block();
            const uint32_t head = bcast0((uint32_t)v_bhead[c]);
            const int h0 = (int)bcast0((uint32_t)(int)heads[c * kHeadsStride + e0]);
            ProbeResult R;
            int f_next, cand_slot;
            bool single_round;
            if (!probe_chain(win, nodes + c * kNodesStride, h0, q, sbuf_len, depth0, depth1, depth2, lane, v_spos, R, f_next, cand_slot,
                             single_round)) {
                if (STATS) nabort++;
                continue;
            }
            // words updates of the two possible outcomes (lz.rs:208,233): words[hash2(s'-3)] = bytes s'-2, s'-1
            const uint32_t hw_lit = (b_m2 & 0x7fu) | ((b_m3 & 0x7fu) | is_alnum(b_m4) << 7) << 7; // s' = q+1
            const uint32_t wv_lit = b_m1 | (cur4 & 0xffu) << 8;
            uint32_t hw_m = 0, wv_m = 0, sym = 0, aux = 0, fmw = 0, bump_pl = 0;
            if (R.fm_len > 0) { // s' = q + len
                const uint32_t t4 = ld_u32_unaligned(win, q + R.fm_len - 5); // bytes s'-5 .. s'-2
                const uint32_t t2 = ld_u32_unaligned(win, q + R.fm_len - 2); // bytes s'-2 .. s'+1
                const uint32_t x3 = (t4 >> 16) & 0xffu, x4 = (t4 >> 8) & 0xffu, x5 = t4 & 0xffu;
                hw_m = (x3 & 0x7fu) | ((x4 & 0x7fu) | is_alnum(x5) << 7) << 7;
                wv_m = t2 & 0xffffu;
                match_fields(head, (uint32_t)R.fm_idx, (uint32_t)R.fm_len, node_lenmin(R.fm_pl), R.fm_exp, sym, aux);
                const uint32_t new_head = head + 1 == kRing ? 0 : head + 1;
                const bool bump = node_lenmin(R.fm_pl) <= (uint32_t)R.fm_len && (uint32_t)R.fm_idx != new_head; // matcher.rs:66-71
                const uint32_t lm = (uint32_t)R.fm_len + 1 < 127u ? (uint32_t)R.fm_len + 1 : 127u;
                bump_pl = node_pos(R.fm_pl) | lm << 25;
                fmw = (uint32_t)R.fm_len | (uint32_t)R.fm_idx << 8 | (bump ? 1u << 20 : 0u);
            }
            __threadfence_block();
            if (bcast0(v_bver[c]) != v0) { // the bucket changed while it was being read: rebuild
                if (STATS) nvfail++;
                continue;
            }
            // ---- publish ----
            {
                // ring-slot image as 32 int16: [0,1] pl = q, [2] expected (patched by the executor), [3] next[0] = h0,
                // [4..31] next[1..28] = the old head's next[0..27]; lane L>=1 holds the old head's next[L-1] in f_next
                const int up = __shfl_sync(kFull, f_next, lane >= 3 ? lane - 3 : 1);
                int hv;
                if (lane == 0) hv = (int)(q & 0xffffu);
                else if (lane == 1) hv = (int)((q & 0x1ffffffu) >> 16);
                else if (lane == 2) hv = 0;
                else if (lane == 3) hv = h0;
                else hv = up;
                reinterpret_cast<int16_t *>(pub.rec)[lane] = (int16_t)hv;
                pub.cand[lane] = (uint16_t)(cand_slot < 0 ? 0xffff : cand_slot);
            }
            if (lane == 0) {
                const uint32_t m1 = R.max_d1 < 0 ? 0u : (R.max_d1 > 255 ? 255u : (uint32_t)R.max_d1);
                const uint32_t m2 = R.max_d2 < 0 ? 0u : (R.max_d2 > 255 ? 255u : (uint32_t)R.max_d2);
                pub.ver = v0;
                pub.cmx = c | m1 << 8 | m2 << 16 | (single_round ? 1u << 24 : 0u);
                pub.hc = h2 | (cur4 & 0xffffu) << 16;
                pub.fm = fmw;
                pub.sym = sym;
                pub.aux = aux;
                pub.bump_pl = bump_pl;
                pub.e0 = e0;
                pub.w_lit = hw_lit | wv_lit << 16;
                pub.w_m = hw_m | wv_m << 16;
                pub.fm_plexp = node_lenmin(R.fm_pl) << 8 | (R.fm_exp & 0xffu);
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) vpub->q = q;
            my_q = q;
            my_c = c;
            my_ver = v0;
            if (STATS) {
                bcyc += clock64() - tb0;
                const uint32_t d = q - bcast0(*v_spos);
                lead[d >= 0x80000000u ? 0 : (d < 3 ? 0 : (d < 8 ? 1 : (d < 16 ? 2 : 3)))]++;
            }
        }
        if (STATS && lane == 0) {
            atomicAdd(&S->dbg[4], nbuild);
            atomicAdd(&S->dbg[5], nabort + (nvfail << 32));
            if (dbg_flags & 32u) { // builder timing report replaces the executor's spin counters
                atomicAdd(&S->dbg[1], bcyc);
                atomicAdd(&S->dbg[2], lead[0] | lead[1] << 32);
                atomicAdd(&S->dbg[3], lead[2] | lead[3] << 32);
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < 32768 / 8; i += NT) reinterpret_cast<uint4 *>(S->words)[i] = reinterpret_cast<const uint4 *>(sh.words)[i];
    for (int i = tid; i < kNumCtx; i += NT) S->ctx_count[i] = sh.cnt[i];
    for (int i = tid; i < kNumBuckets; i += NT) S->bhead[i] = sh.bhead[i];
}

} // namespace orz
#include "parse_v4.cuh"
#include "parse_v7.cuh"
namespace orz {

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
    uint32_t stage[kEmitWarps][32];   // stage B: a batch's (symbol | unlikely << 16) for the lane that runs the rank coder
    uint16_t rank_out[kEmitWarps][32];
    uint32_t rcp[kSymStride];         // ceil(2^32 / d): exact (esum >> 4) / d by multiply-high
    uint32_t next_ctx;
    uint32_t num_counted;
    uint32_t hdr_bits;
    uint32_t total_bits;
};

__global__ void __launch_bounds__(kEmitThreads) orz_emit_kernel(EncStream *streams, int buf) {
    EncStream *S = &streams[blockIdx.x];
    if (S->status != 0) return;
    if (S->v7 && S->self_driven && S->chunk_flags[buf] == 2u) return; // nothing was parsed into this buffer
    ORZ_DYN_SHARED(smem_raw);
    EmitShared &sh = *reinterpret_cast<EmitShared *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Item *items = buf ? S->items_alt : S->items;
    uint32_t *sorted = S->sorted;
    const uint32_t n_items = buf ? S->n_items_alt : S->n_items;
    const bool first_block = S->first_block != 0;

    // ---- init shared ----
    for (int i = tid; i < 2 * kSymStride; i += kEmitThreads) (&sh.hist1[0][0])[i] = 0;
    for (int i = tid; i < 256; i += kEmitThreads) sh.hist2[i] = 0;
    for (int i = tid; i < kSymStride; i += kEmitThreads) sh.symcount[i] = 0;
    for (int i = tid; i < kNumCtx; i += kEmitThreads) sh.cnt[i] = (buf ? S->ctx_count_alt : S->ctx_count)[i];
    for (int i = tid; i < kSymStride; i += kEmitThreads) sh.rcp[i] = i < 2 ? 0u : (uint32_t)(((1ull << 32) + (unsigned)i - 1) / (unsigned)i);
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
                // The rank update is a serial state machine per context (symrank.rs:38-97), and the largest context of a chunk
                // (about a tenth of its items) is the critical path of this kernel.  So ONE lane runs it on the shared table
                // with nothing else on its dependency chain: the batch's symbols are staged in shared memory by all lanes, the
                // ranks come back the same way, and sum/16/cnt is a multiply-high with a precomputed reciprocal.
                sh.stage[warp][lane] = sym | unl << 16;
                __syncwarp();
                if (lane == 0) {
                    for (int q = 0; q < m; q++) {
                        const uint32_t w = sh.stage[warp][q];
                        const uint32_t v = w & 0xffffu, vu = w >> 16;
                        const uint32_t i = tab.index[v];
                        const uint32_t iun = tab.index[vu];
                        if (ecnt > (uint32_t)kNumSym) { // symrank.rs:61-66
                            ecnt = ecnt * 9 / 10;
                            esum = esum * 9 / 10;
                        }
                        ecnt += 1;
                        esum += i;
                        // (esum >> 4) / ecnt: exact by reciprocal while (esum >> 4) * ecnt < 2^32 (esum starts at 1,000,000 and decays)
                        const uint32_t a16 = esum >> 4;
                        const uint32_t quo = (ecnt < 2 || esum >= (1u << 24)) ? a16 / ecnt : __umulhi(a16, sh.rcp[ecnt]);
                        const uint32_t sub = ((i >> 4) + quo) & 0xffffu; // u16 arithmetic
                        uint32_t next_i = i > sub ? i - sub : 0;
                        if (next_i < (i >> 1)) next_i = i >> 1;
                        const uint32_t d = i - next_i;
                        if (d == 1) { // symrank.rs:75-97: 2- or 3-way rotation
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
                        sh.rank_out[warp][q] = (uint16_t)((i == iun) ? (uint32_t)(kNumSym - 1) : i - (i > iun)); // symrank.rs:43-46
                    }
                }
                __syncwarp();
                const uint32_t myrank = sh.rank_out[warp][lane];
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
        // (v7: stage A of the next chunk may already be moving spos; the chunk's own end was recorded by orz_chunk_done_kernel)
        const uint32_t end_spos = S->v7 ? S->chunk_end_spos[buf] : (S->spos < S->sbuf_len ? S->spos : S->sbuf_len);
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
    if (!fits) {
        if (tid == 0) S->status = -2; // ORZ_ENOSPC
        return;
    }
    if (S->v7) { // the loop state moved on with stage A (orz_chunk_done_kernel + advance kernels); only the end marker is left
        if (tid == 0) {
            if (S->chunk_flags[buf] & 1u) { // lib.rs:89: empty chunk marks EOF
                S->out[opos] = 0;
                S->out_pos = opos + 1;
                S->status = 1;
            } else {
                S->out_pos = opos;
            }
        }
        return;
    }
    const uint32_t spos = S->spos, sbuf_len = S->sbuf_len;
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

// After stage A of a chunk (v7 streams, one thread per stream): records what stages B-D need of the loop state and takes the
// lib.rs:72-89 decision — next chunk, end of the block (the advance kernels follow) or end of the stream — so that stage A of
// the next chunk does not wait for stages B-D of this one.
__global__ void orz_chunk_done_kernel(EncStream *streams, int nstreams, int buf) {
    const int si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= nstreams) return;
    EncStream *S = &streams[si];
    if (S->status != 0 || !S->v7) return;
    if (S->parse_done) { // an earlier chunk was the stream's last: nothing was parsed into this buffer
        S->chunk_flags[buf] = 2u;
        return;
    }
    const uint32_t spos = S->spos, sbuf_len = S->sbuf_len;
    S->chunk_end_spos[buf] = spos < sbuf_len ? spos : sbuf_len;
    uint32_t flags = 0;
    if (S->self_driven && spos >= sbuf_len) {
        if (S->in_pos >= S->in_len) {
            flags = 1u;
            S->parse_done = 1;
        } else {
            S->need_slide = 1;
        }
    }
    S->chunk_flags[buf] = flags;
}

// End of a block for v7 streams (lib.rs:83-84,72), grid-wide.  phase 0: slide the window (copy_within(len-PREMATCH.., 0)) and
// rebase the key lists (LZEncoder::forward, lz.rs:82-87); phase 1: read the next block into [PREMATCH, PREMATCH+rd).
// The two phases are separate launches because the load overwrites the slide's source.
__global__ void __launch_bounds__(256) orz_enc_advance_kernel(EncStream *streams, int nstreams, int phase) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
    for (int si = 0; si < nstreams; si++) {
        EncStream *S = &streams[si];
        if (S->status != 0 || !S->need_slide) continue;
        uint8_t *win = S->win;
        if (phase == 0) {
            for (size_t i = gid; i < (kPrematch + 15) / 16; i += gsz) {
                // source and destination are 16,777,216 apart and 16 B aligned; ranges do not overlap
                const uint4 v = *reinterpret_cast<const uint4 *>(win + kBlockNew + 16u * i);
                if (16u * i + 16u <= kPrematch) {
                    *reinterpret_cast<uint4 *>(win + 16u * i) = v;
                } else {
                    const uint8_t *vb = reinterpret_cast<const uint8_t *>(&v);
                    for (uint32_t b = 0; 16u * i + b < kPrematch; b++) win[16u * i + b] = vb[b];
                }
            }
            enc7_forward_rebase(S->klist, S->kcnt, kBlockNew, gid, gsz);
        } else {
            const uint64_t in_pos = S->in_pos, in_len = S->in_len;
            const uint64_t rem = in_len - in_pos;
            const uint32_t rd = rem < kBlockNew ? (uint32_t)rem : kBlockNew;
            const uint8_t *src = S->in + in_pos;
            // win + PREMATCH is 1 short of 16-byte alignment: byte head, then 16-byte stores from 16-byte loads when `src` allows
            const uint32_t headb = rd < 1u ? rd : 1u;
            if (gid == 0) for (uint32_t i = 0; i < headb; i++) win[kPrematch + i] = src[i];
            const uint32_t body = (rd - headb) / 16u;
            if (((uintptr_t)(src + headb) & 15u) == 0) {
                for (size_t i = gid; i < body; i += gsz)
                    *reinterpret_cast<uint4 *>(win + kPrematch + headb + 16u * i) = *reinterpret_cast<const uint4 *>(src + headb + 16u * i);
            } else {
                for (size_t i = gid; i < (size_t)body * 16u; i += gsz) win[kPrematch + headb + i] = src[headb + i];
            }
            for (size_t i = (size_t)headb + (size_t)body * 16u + gid; i < rd; i += gsz) win[kPrematch + i] = src[i];
        }
    }
}
// after phase 1 of every stream: the lib.rs:72 loop state of the new block (one thread per stream)
__global__ void orz_enc_advance_done_kernel(EncStream *streams, int nstreams) {
    const int si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= nstreams) return;
    EncStream *S = &streams[si];
    if (S->status != 0 || !S->need_slide) return;
    const uint64_t in_pos = S->in_pos, rem = S->in_len - in_pos;
    const uint32_t rd = rem < kBlockNew ? (uint32_t)rem : kBlockNew;
    S->in_pos = in_pos + rd;
    S->spos = kPrematch;
    S->sbuf_len = kPrematch + rd;
    S->need_slide = 0;
}

// LZEncoder::forward for the chunk-level API (lz.rs:82-87), v7 layout
__global__ void __launch_bounds__(256) orz_enc7_forward_kernel(EncStream *streams, uint32_t forward_len) {
    EncStream *S = &streams[0];
    enc7_forward_rebase(S->klist, S->kcnt, forward_len, (size_t)blockIdx.x * blockDim.x + threadIdx.x, (size_t)gridDim.x * blockDim.x);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is per device: set it before every launch (cheap), not once per process
template <class Kern>
static void set_smem_attr(Kern kern, size_t bytes) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <int NW, int NB>
static void launch_parse_variant(EncStream *d_streams, int n, cudaStream_t st) {
    auto kern = orz_parse_kernel<NW, NB, false>;
    set_smem_attr(kern, sizeof(ParseShared));
    ORZ_LAUNCH(kern, n, NW * 32, sizeof(ParseShared), st, d_streams);
}

template <int NW, int SL>
static void launch_parse4_variant(EncStream *d_streams, int n, cudaStream_t st) {
    typedef ParseShared4<NW - 1, SL> Sh;
    static_assert(sizeof(Sh) <= 227 * 1024, "stage A v4: shared memory exceeds an SM");
    auto kern = orz_parse4_kernel<NW, SL, false>;
    set_smem_attr(kern, sizeof(Sh));
    ORZ_LAUNCH(kern, n, NW * 32, sizeof(Sh), st, d_streams);
}

template <int NW, int K, bool STATS>
static void launch_parse7_variant(EncStream *d_streams, int n, int buf, cudaStream_t st) {
    static_assert(sizeof(ParseShared7) <= 227 * 1024, "stage A v7: shared memory of CTA 0 exceeds an SM");
    auto kern = orz_parse7_kernel<NW, K, STATS>;
    set_smem_attr(kern, sizeof(ParseShared7));
#ifndef ORZ_SIMT_EMU
    if (K > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
#endif
    if (K > 1) ORZ_LAUNCH_CLUSTER(kern, n * K, NW * 32, sizeof(ParseShared7), st, K, d_streams, buf);
    else ORZ_LAUNCH(kern, n, NW * 32, sizeof(ParseShared7), st, d_streams, buf);
}

// Library configuration, read once per process:
//   ORZ_B200_STATS=1            stage A v7 runs its instrumented instantiation (counters + clock64 phases, printed by capi.cu)
//   ORZ_B200_V7_CLUSTER=1|8     CTAs per stream of stage A v7 (default 8; 1 = builders beside the resolver, the emulator tests' configuration)
//   ORZ_B200_LEGACY_STAGE_A=1|v3|v3s|v4   ring-record layout instead of key lists (capi.cu reads it when a batch is created);
//                               v3 / v3s / v4 additionally pin the kernel (tests)
struct LibCfg {
    bool stats;
    int cluster;
    int legacy; // 0 automatic, 3 = v3<22,16>, 13 = v3<11,8>, 4 = v4<16,4>
    LibCfg() {
        stats = getenv("ORZ_B200_STATS") != nullptr;
        const char *e = getenv("ORZ_B200_V7_CLUSTER");
        cluster = e ? atoi(e) : 8;
        e = getenv("ORZ_B200_LEGACY_STAGE_A");
        legacy = 0;
        if (e && !strcmp(e, "v3")) legacy = 3;
        else if (e && !strcmp(e, "v3s")) legacy = 13;
        else if (e && !strcmp(e, "v4")) legacy = 4;
    }
};
static const LibCfg &lib_cfg() {
    static const LibCfg c;
    return c;
}
bool parse7_stats_enabled() { return lib_cfg().stats; }

// Stage A.  v7: the stream arena holds key lists (parse_v7.cuh: one cluster per stream).  Otherwise ring records +
// hash chains: one stream runs v4 (windowed resolver, 16 warps), several streams run the v3 kernel (one executor thread + 8 or
// 16 builder warps per stream), whose many-stream throughput is the measured one.
void launch_parse(EncStream *d_streams, int n, bool v7, int buf, cudaStream_t st) {
    const LibCfg &c = lib_cfg();
    if (v7) {
        if (c.stats) {
            if (c.cluster == 1) launch_parse7_variant<16, 1, true>(d_streams, n, buf, st);
            else launch_parse7_variant<16, 8, true>(d_streams, n, buf, st);
        } else {
            if (c.cluster == 1) launch_parse7_variant<16, 1, false>(d_streams, n, buf, st);
            else launch_parse7_variant<16, 8, false>(d_streams, n, buf, st);
        }
        return;
    }
    if (c.legacy == 4 || (c.legacy == 0 && n == 1)) {
        launch_parse4_variant<16, 4>(d_streams, n, st);
        return;
    }
    // few streams: 16 builders hide more latency per stream; many streams: 8 builders, more streams per SM
    if (c.legacy == 3 || (c.legacy == 0 && n <= 2 * 148)) launch_parse_variant<22, 16>(d_streams, n, st);
    else launch_parse_variant<11, 8>(d_streams, n, st);
}
void launch_emit(EncStream *d_streams, int n, int buf, cudaStream_t st) {
    set_smem_attr(orz_emit_kernel, sizeof(EmitShared));
    ORZ_LAUNCH(orz_emit_kernel, n, kEmitThreads, sizeof(EmitShared), st, d_streams, buf);
}
void launch_chunk_done(EncStream *d_streams, int n, int buf, cudaStream_t st) { ORZ_LAUNCH(orz_chunk_done_kernel, (n + 63) / 64, 64, 0, st, d_streams, n, buf); }
void launch_enc_advance(EncStream *d_streams, int n, cudaStream_t st) {
    ORZ_LAUNCH(orz_enc_advance_kernel, 592, 256, 0, st, d_streams, n, 0);
    ORZ_LAUNCH(orz_enc_advance_kernel, 592, 256, 0, st, d_streams, n, 1);
    ORZ_LAUNCH(orz_enc_advance_done_kernel, (n + 63) / 64, 64, 0, st, d_streams, n);
}
void launch_enc_begin(EncStream *d_streams, int n, cudaStream_t st) { ORZ_LAUNCH(orz_enc_begin_kernel, n, 256, 0, st, d_streams); }
void launch_enc_forward(EncStream *d_stream, uint32_t forward_len, bool v7, cudaStream_t st) {
    if (v7) {
        ORZ_LAUNCH(orz_enc7_forward_kernel, 592, 256, 0, st, d_stream, forward_len);
        return;
    }
    ORZ_LAUNCH(orz_enc_forward_kernel, 148, 256, 0, st, d_stream, forward_len, 0);
    ORZ_LAUNCH(orz_enc_forward_kernel, 148, 256, 0, st, d_stream, forward_len, 1);
}

} // namespace orz
