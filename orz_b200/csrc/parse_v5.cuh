// parse_v5.cuh — stage A, v4 generalised (included by encode_kernels.cu; opt-in with ORZ_B200_PARSE=5).
//
// parse_v4.cuh is the B200-verified single-CTA kernel and stays untouched; this file is the same resolver and the same
// lock-free builders with three compile-time extensions, written and checked in the SIMT emulator in round 1 (the GPU
// budget of the round was spent) and to be measured on the GPU in round 2:
//
//   K    > 1 : a thread-block CLUSTER of K CTAs per stream.  CTA 0 holds the resolver, the mailbox, the per-bucket
//              logs and `words` in its shared memory exactly as v4; the warps of CTAs 1..K-1 are all builders.  They
//              read spos / bcnt and publish their probes straight into CTA 0's shared memory over DSMEM
//              (map_shared_rank), so the resolver's window load stays a shared-memory load.  v4 on a B200 is bound by
//              builder throughput (4.3 builds of ~4500 cycles per step over 15 builders); a cluster of 8 has 24 + 7*32
//              builders.  Remote builders read the ring records and heads with ld.global.cg (their SM's L1 is not
//              coherent with the resolver's stores) and all publication fences are gpu-scope.
//   PARK     : in CTA 0 the warps that share the resolver's SM sub-partition (warp % 4 == 0) are parked at the final
//              barrier, as in v3: a spinning builder is always ready to issue and takes issue slots from the resolver
//              (measured: v4 with 32 warps, 7 builders beside the resolver, is SLOWER than with 16).
//   idle builders sleep (__nanosleep) instead of spinning on the mailbox.
#pragma once
#ifndef ORZ_SIMT_EMU
#include <cooperative_groups.h>
#endif

namespace orz {

constexpr int kBlog5 = 16; // inserts remembered per bucket (the look-ahead of a cluster is deeper than v4's)
// Builders do not run further ahead of the parse than this many positions (~40 steps): a probe must be ready one build
// latency before its turn, and every step it waits beyond that is a chance to go stale (3.5 % within 32 steps, 5.4 %
// within 64; the per-bucket log holds 16 inserts).  With more builders than that needs, the surplus sleeps.
constexpr uint32_t kLookahead5 = 192;

#ifdef ORZ_SIMT_EMU
template <class T>
__device__ __forceinline__ T *map_to_cta0(T *p) { return reinterpret_cast<T *>(simt::cur->cluster_smem[0] + (reinterpret_cast<uint8_t *>(p) - simt::cur->dyn_smem)); }
__device__ __forceinline__ void cluster_barrier() { simt::cluster_sync(); }
#else
template <class T>
__device__ __forceinline__ T *map_to_cta0(T *p) { return cooperative_groups::this_cluster().map_shared_rank(p, 0); }
__device__ __forceinline__ void cluster_barrier() { cooperative_groups::this_cluster().sync(); }
#endif

// ring records / heads: through L1 for builders on the resolver's SM, L2-only (ld.global.cg) for remote builders
template <bool CG>
__device__ __forceinline__ uint32_t ld_ring(const uint32_t *p) { return CG ? __ldcg(p) : *p; }

template <bool CG>
__device__ __forceinline__ void enode_head5(const ENode *bn, int slot, uint32_t &pl, uint32_t &exp, int &my_next, int lane) {
    const uint32_t *r = reinterpret_cast<const uint32_t *>(bn + slot);
    pl = __shfl_sync(0xffffffffu, ld_ring<CG>(r), 0);
    exp = __shfl_sync(0xffffffffu, ld_ring<CG>(r + 1), 0) & 0xffu;
    my_next = -1;
    if (lane >= 1 && lane <= kSnap) {
        const uint32_t w = ld_ring<CG>(r + ((lane + 2) >> 1)); // int16 index lane+2 of the record
        my_next = (int)(int16_t)(((lane + 2) & 1) ? (w >> 16) : (w & 0xffffu));
    }
}

template <bool CG>
__device__ __forceinline__ int gather_round5(const ENode *bn, int lane, int cap, uint32_t x_pl, uint32_t x_exp, int x_idx, int my_next,
                                             uint32_t &my_pl, uint32_t &my_exp, int &my_idx) {
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
        my_pl = ld_ring<CG>(rr);
        my_exp = ld_ring<CG>(rr + 1) & 0xffu;
        my_idx = my_next;
        ok = true;
    }
    const uint32_t prev_pos = __shfl_up_sync(kFull, node_pos(my_pl), 1);
    const bool link_ok = (lane == 0) || (ok && node_pos(my_pl) < prev_pos);
    const unsigned bad = ~__ballot_sync(kFull, link_ok);
    return bad ? (__ffs(bad) - 1) : 32;
}

template <bool GPU>
__device__ __forceinline__ void fence5() {
    if (GPU) __threadfence();
    else __threadfence_block();
}
template <bool CG>
__device__ __forceinline__ int ld_head(const int16_t *p) {
#ifndef ORZ_SIMT_EMU
    if (CG) return (int)__ldcg(reinterpret_cast<const short *>(p));
#endif
    return (int)*p;
}

// probe_chain4 with selectable load flavour; probe_chain of v3 plus the distance bookkeeping: `safe` = how many further inserts into the bucket (ring head
// `head` when the build began) leave every ring slot alone that this walk read (all rounds, the stop slot included).
template <bool CG>
__device__ __forceinline__ bool probe_chain5(const uint8_t *win, const ENode *bn, int h0, uint32_t q, uint32_t buf_len, int depth0,
                                             int depth1, int depth2, int lane, const volatile uint32_t *s_spos, uint32_t head,
                                             ProbeResult &R, int &f_next, uint32_t &safe) {
    R.fm_len = 0;
    R.fm_idx = 0;
    R.fm_pl = 0;
    R.fm_exp = 0;
    R.max_d1 = -1;
    R.max_d2 = -1;
    f_next = -1;
    safe = 0xffffu;
    if (h0 < 0) return true;
    uint32_t x_pl, x_exp;
    int x_idx = h0, my_next;
    enode_head5<CG>(bn, h0, x_pl, x_exp, my_next, lane);
    ORZ_COST(450); // head record: one L2 round trip
    f_next = my_next;
    int dmax = depth0 > depth1 ? depth0 : depth1;
    dmax = dmax > depth2 ? dmax : depth2;
    uint32_t my_dist = 0xffffu;
    {
        const uint32_t d = ((uint32_t)h0 + kRing - head) % kRing;
        if (lane == 0) my_dist = d == 0 ? (uint32_t)kRing : d;
    }
    if (dmax > 0) {
        int max_len = kMatchMin - 1;
        uint32_t exp_best = kMatchMin, pl_best = (uint32_t)kMatchMin << 25;
        int idx_best = 0;
        int seen = 0;
        bool fm_done = depth0 <= 0;
        bool first_round = true;
        for (;;) {
            if (bcast0(ld_vol_u32(s_spos)) > q) return false;
            const int lo = first_round ? 0 : 1;
            const int room = dmax - seen - (first_round ? 1 : 0);
            const int cap = room < kSnap ? room : kSnap;
            uint32_t my_pl, my_exp;
            int my_idx;
            const int nvl = gather_round5<CG>(bn, lane, cap, x_pl, x_exp, x_idx, my_next, my_pl, my_exp, my_idx);
            ORZ_COST(450); // candidate records: one L2 round trip for the whole gather
            if (lane <= nvl && my_idx >= 0) { // every record that decided something, the one the walk stopped at included
                const uint32_t d = ((uint32_t)my_idx + kRing - head) % kRing;
                const uint32_t dd = d == 0 ? (uint32_t)kRing : d;
                my_dist = dd < my_dist ? dd : my_dist;
            }
            const int gi = seen + lane - lo;
            const bool have = lane >= lo && lane < nvl;
            int lcp = -1;
            if (have) lcp = lcp_win_win(win, node_pos(my_pl), q);
            ORZ_COST(500); // candidate bytes: L2 round trip + compare loop

            R.max_d1 = max(R.max_d1, warp_max((have && gi < depth1) ? lcp : -1));
            R.max_d2 = max(R.max_d2, warp_max((have && gi < depth2) ? lcp : -1));

            if (!fm_done) { // matcher.rs:149-181, as v3
                const bool valid = have && gi < depth0;
                int cur = lo - 1;
                for (;;) {
                    const unsigned after = (cur >= 31) ? 0u : (0xffffffffu << (cur + 1));
                    const unsigned upd = __ballot_sync(kFull, valid && lcp > max_len) & after;
                    const int u = upd ? (__ffs(upd) - 1) : 32;
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
                    if (max_len == kMatchMax || (exp_best > 0 && max_len > (int)exp_best)) {
                        fm_done = true;
                        break;
                    }
                    cur = u;
                }
            }
            seen += nvl - lo;
            if (seen >= depth0) fm_done = true;
            const bool lazy_done = (seen >= depth1 || R.max_d1 >= kMatchMax / 2 + 2) && (seen >= depth2 || R.max_d2 >= kMatchMax / 2 + 2);
            if (nvl <= cap || (fm_done && lazy_done)) break;
            x_pl = __shfl_sync(kFull, my_pl, kSnap);
            x_exp = __shfl_sync(kFull, my_exp, kSnap);
            x_idx = __shfl_sync(kFull, my_idx, kSnap);
            uint32_t d0, d1;
            enode_head5<CG>(bn, x_idx, d0, d1, my_next, lane);
            ORZ_COST(450);
            first_round = false;
        }
        if (max_len >= kMatchMin && q + max_len < buf_len) { // matcher.rs:183
            R.fm_len = max_len;
            R.fm_idx = idx_best;
            R.fm_pl = pl_best;
            R.fm_exp = exp_best;
        }
    }
    // warp minimum of the distances; `safe` = that many inserts minus one are harmless
    for (int o = 16; o; o >>= 1) {
        const uint32_t other = __shfl_xor_sync(kFull, my_dist, o);
        my_dist = other < my_dist ? other : my_dist;
    }
    safe = my_dist - 1;
    return true;
}

template <int NB, int SL>
struct ParseShared5 {
    uint16_t words[32768];          // lz.rs:52 — only the resolver touches it
    uint32_t cnt[kNumCtx];
    uint32_t bcnt[kNumBuckets];     // inserts into the bucket so far (ring head = bcnt % 4094)
    uint32_t blog[kNumBuckets][kBlog5]; // the last inserts of the bucket: e0 | bumped slot << 16 (0xfff = none)
    uint32_t spos;                  // published parse position
    uint32_t done;
    uint32_t pad[2];
    Probe4 pub[NB * SL];            // mailbox: position q lives in slot (q % NB) * SL + (q / NB) % SL
};

// K CTAs of NW warps per stream (a cluster when K > 1).  CTA 0: warp 0 is the resolver, the other warps build
// (PARK: except those on the resolver's SM sub-partition); CTAs 1..K-1: every warp builds.  SL mailbox slots per
// builder (look-ahead = NB * SL positions).
template <int NW, int SL, int K, bool PARK>
struct Parse5Dims {
    static constexpr int NB_LOCAL = PARK ? NW - 1 - (NW - 1) / 4 : NW - 1;
    static constexpr int NB = NB_LOCAL + (K - 1) * NW;
};

template <int NW, int SL, int K, bool PARK, bool STATS>
__global__ void __launch_bounds__(NW * 32) orz_parse5_kernel(EncStream *streams) {
    constexpr int NB_LOCAL = Parse5Dims<NW, SL, K, PARK>::NB_LOCAL;
    constexpr int NB = Parse5Dims<NW, SL, K, PARK>::NB;
    constexpr int NT = NW * 32;
    constexpr bool CL = K > 1;
    const int rank = CL ? (int)(blockIdx.x % K) : 0;
    EncStream *S = &streams[blockIdx.x / K];
    // every CTA of a cluster takes the same branch (the status word is not written before the final cluster barrier)
    if (S->status != 0) return;
    ORZ_DYN_SHARED(smem_raw);
    ParseShared5<NB, SL> *sh_local = reinterpret_cast<ParseShared5<NB, SL> *>(smem_raw);
    ParseShared5<NB, SL> &sh = CL ? *map_to_cta0(sh_local) : *sh_local; // the resolver CTA's shared memory (DSMEM for rank > 0)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (rank == 0) {
        for (int i = tid; i < kNumCtx; i += NT) sh.cnt[i] = 0;
        for (int i = tid; i < kNumBuckets; i += NT) sh.bcnt[i] = S->bhead[i];
        for (int i = tid; i < kNumBuckets * kBlog5; i += NT) (&sh.blog[0][0])[i] = 0xffffffffu;
        for (int i = tid; i < 32768 / 8; i += NT) reinterpret_cast<uint4 *>(sh.words)[i] = reinterpret_cast<const uint4 *>(S->words)[i];
        for (int i = tid; i < NB * SL; i += NT) sh.pub[i].q = kNone;
    }
    const uint8_t *win = S->win;
    ENode *nodes = S->nodes;
    int16_t *heads = S->heads;
    const uint32_t sbuf_len = S->sbuf_len;
    const int depth0 = (int)S->depth0, depth1 = (int)S->depth1, depth2 = (int)S->depth2;
    if (rank == 0 && tid == 0) {
        sh.spos = S->spos;
        sh.done = 0;
    }
    __syncthreads();
    if (CL) cluster_barrier(); // CTA 0's shared memory is initialised before any remote builder looks at it
    volatile uint32_t *v_spos = &sh.spos;
    volatile uint32_t *v_done = &sh.done;
    volatile uint32_t *v_bcnt = sh.bcnt;

    if (rank == 0 && warp == 0) {
        // ------------------------------------------------------------------ resolver (one warp, lane l = position spos + l)
        Item *items = S->items;
        const uint32_t items_cap = S->items_cap < kChunkItems ? S->items_cap : kChunkItems;
        uint32_t spos = S->spos;
        uint32_t after_literal = S->after_literal;
        uint32_t n_items = 0;
        unsigned long long n_windows = 0, n_empty = 0, n_inval = 0;
        unsigned idle = 0;
        int wd_code = 0;
        const unsigned lt_mask = (1u << lane) - 1u;
        while (spos < sbuf_len && n_items < items_cap) {
            const uint32_t p = spos + (uint32_t)lane;
            // ---- load and validate this lane's probe against everything committed so far
            Probe4 *P = &sh.pub[(p % NB) * SL + (p / NB) % SL];
            const uint4 a = lds_volatile_v4(P);
            const uint4 b = lds_volatile_v4(&P->mx);
            const uint32_t q2 = ld_vol_u32(&P->q);
            const bool ready = p < sbuf_len && (a.x & kPosMask) == p && a.x != kNone && a.x == q2;
            const uint32_t c = a.z & 0xffu, e = (a.z >> 8) & 0x1fffu;
            const uint32_t fm_len = a.w & 0xffu, fm_idx = (a.w >> 8) & 0xfffu;
            const uint32_t head0 = sh.bcnt[c];            // committed inserts into this lane's bucket
            const uint32_t dnew = head0 - a.y;            // ... of which this many are younger than the probe
            const uint32_t safe = a.z >> 21;
            bool ok = ready && dnew <= (uint32_t)kBlog5 && dnew <= safe;
            if (ok) {
                for (uint32_t k = 0; k < dnew; k++) {
                    const uint32_t ent = sh.blog[c][(a.y + k) % kBlog5];
                    if ((ent & 0xffffu) == e || (fm_len > 0 && (ent >> 16) == fm_idx)) ok = false;
                }
            }
            if (ready && !ok) { // stale: hand the slot back to its builder (unless it is already rebuilding it)
                atomicCAS(&P->q, a.x, kNone);
                if (STATS) n_inval++;
            }
            uint32_t room0 = ok ? safe - dnew : 0u;       // further inserts into c that leave the probe's slots alone
            // ---- neighbours' lazy information (lane+1: depth1 maximum, lane+2: depth2 maximum)
            const uint32_t st = (ok ? 1u : 0u) | room0 << 1 | c << 8 | e << 16;      // e: 13 bits, room: 6 bits
            const uint32_t st1 = __shfl_down_sync(kFull, st, 1), st2 = __shfl_down_sync(kFull, st, 2);
            const uint32_t mx1 = __shfl_down_sync(kFull, b.x, 1), mx2 = __shfl_down_sync(kFull, b.x, 2);
            bool ok1 = (st1 & 1u) && lane < 31, ok2 = (st2 & 1u) && lane < 30;
            uint32_t room1 = (st1 >> 1) & 0x3fu, room2 = (st2 >> 1) & 0x3fu;
            const uint32_t c1 = (st1 >> 8) & 0xffu, c2 = (st2 >> 8) & 0xffu, e1 = st1 >> 16, e2 = st2 >> 16;
            const uint32_t md1 = mx1 & 0xffu, md2 = (mx2 >> 8) & 0xffu;
            const uint32_t h2 = b.y & 0x7fffu, cur2 = b.y >> 16;
            uint32_t lwe = sh.words[h2];                  // lz.rs:132 (kept current below while the window advances)
            bool conf0 = !ok, conf1 = !ok1, conf2 = !ok2;
            bool stale_ce = false; // a step of this window inserted this probe's own hash entry: stale for certain once committed
            // what this lane decided when its turn came
            uint32_t my_kind = 0, my_head = 0, my_al = 0, my_lwe = 0;
            unsigned stepmask = 0;
            uint32_t cur = 0, n_new = 0;

            // ---- walk the step chain through the window (lz.rs:131-236); every value below is lane-local
            // (the loop is the serial chain of the stream: no division, no packing beyond one word, all four broadcasts issued together)
            // Everything that does not change while the window advances is folded into flags here, so that a lane's decision
            // is a handful of selects on (rlt, lwm, conf*): lazy thresholds are fm_len+1+rlt and that minus lwm (lz.rs:151-153).
            const bool need_lazy = fm_len > 0 && fm_len < (uint32_t)(kMatchMax / 2);     // lz.rs:151
            const bool t1a = md1 >= fm_len + 1, t1b = md1 >= fm_len + 2;
            const bool u0 = md2 >= fm_len, u1 = md2 >= fm_len + 1, u2 = md2 >= fm_len + 2;
            const bool can_word = p + 1 < sbuf_len;
            const uint32_t ce = c | e << 8;
            uint32_t head = head0 % kRing;                                               // ring head when this lane's step runs
            bool rlt;                                                                    // robitlen < 8  <=>  reduced offset + 2 < 512
            {
                uint32_t ro = head + kRing - fm_idx;                                     // matcher.rs:185
                ro = ro >= (uint32_t)kRing ? ro - kRing : ro;
                rlt = ro < 510u;
            }
            for (;;) {
                const bool lwm = cur2 == lwe;                                           // lz.rs:133
                int lazy_id = 0;
                bool usable = !conf0;
                if (need_lazy) {
                    if (conf1) usable = false;
                    else if (rlt ? t1b : t1a) lazy_id = 1;
                    else if (conf2) usable = false;
                    else if (rlt == lwm ? u1 : (rlt ? u2 : u0)) lazy_id = 2;
                }
                const bool emit_match = fm_len > 0 && lazy_id == 0;
                const bool word = !emit_match && can_word && lazy_id != 1 && lwm;           // lz.rs:218
                const uint32_t adv = emit_match ? fm_len : (word ? 2u : 1u);
                const uint32_t kind = emit_match ? 2u : (word ? 1u : 0u);
                // the broadcast of this lane's step, should it be the next one
                const uint32_t dec = (usable ? 1u : 0u) | kind << 1 | adv << 3;
                const uint32_t wu = word ? 0xffffffffu : (emit_match ? b.w : b.z);
                // A match step may bump the lenmin of its matched node (matcher.rs:66-71; decided with fm_pl at commit).
                // Until it is committed, every later probe that matched the same node counts as stale.
                const uint32_t d_s = __shfl_sync(kFull, dec, cur), ce_s = __shfl_sync(kFull, ce, cur);
                const uint32_t wu_s = __shfl_sync(kFull, wu, cur), bump_s = __shfl_sync(kFull, emit_match ? fm_idx : kNoBump, cur);
                if (!(d_s & 1u)) break;                                                  // probe (or lazy probe) missing / stale
                const uint32_t c_s = ce_s & 0xffu, e_s = ce_s >> 8, kind_s = (d_s >> 1) & 3u, adv_s = d_s >> 3;
                if ((uint32_t)lane == cur) {
                    my_kind = kind;
                    my_head = head;
                    my_al = after_literal;
                    my_lwe = lwe;
                }
                stepmask |= 1u << cur;
                n_new++;
                after_literal = kind_s == 0 ? 1u : 0u;
                // keep every lane's inputs current: the insert of step `cur` into bucket c_s, and its words update
                if (c == c_s) {
                    if (e == e_s || (fm_len > 0 && fm_idx == bump_s)) conf0 = true;
                    if (e == e_s) stale_ce = true;
                    if (room0 == 0) conf0 = true;
                    else room0--;
                    head = head + 1 == (uint32_t)kRing ? 0u : head + 1;                  // one more insert ahead of this lane's turn
                    uint32_t ro = head + kRing - fm_idx;
                    ro = ro >= (uint32_t)kRing ? ro - kRing : ro;
                    rlt = ro < 510u;
                }
                if (c1 == c_s) {
                    if (e1 == e_s || room1 == 0) conf1 = true;
                    else room1--;
                }
                if (c2 == c_s) {
                    if (e2 == e_s || room2 == 0) conf2 = true;
                    else room2--;
                }
                if (kind_s != 1) { // lz.rs:208,233
                    if (h2 == (wu_s & 0x7fffu)) lwe = wu_s >> 16;
                    if ((uint32_t)lane == cur) sh.words[wu_s & 0x7fffu] = (uint16_t)(wu_s >> 16);
                }
                cur += adv_s;
                if (cur >= 32 || n_items + n_new >= items_cap) break;
            }
            if (STATS) {
                n_windows++;
                if (!stepmask) n_empty++;
            }
            if (!stepmask) { // nothing usable yet: poll again
                if (++idle > 40000000u) {
                    wd_code = 1;
                    break;
                }
                continue;
            }
            idle = 0;

            // ---- commit the window's steps in parallel (matcher.rs:62-80,115-121; lz.rs:173-236)
            const bool is_step = (stepmask >> lane) & 1u;
            const uint32_t ctx = c | my_al << 8;                                         // lz.rs:134-135
            // ordinal of the item among the chunk's items of its context
            const unsigned same_ctx = __match_any_sync(kFull, is_step ? ctx : (0x10000u | (uint32_t)lane));
            const unsigned same_c = __match_any_sync(kFull, is_step ? c : (0x10000u | (uint32_t)lane));
            const uint32_t cnt_base = is_step ? sh.cnt[ctx] : 0u;
            __syncwarp(); // every lane has read its counter before the group leaders advance them
            if (is_step) {
                const uint32_t seq = cnt_base + (uint32_t)__popc(same_ctx & lt_mask);
                const uint32_t new_head = my_head + 1 == (uint32_t)kRing ? 0u : my_head + 1;
                const uint4 v2 = lds_volatile_v4(&P->fm_pl);
                const uint32_t fm_pl = v2.x, fm_exp = a.w >> 20;
                Item it;
                it.ctx = (uint16_t)ctx;
                it.unlikely = (uint8_t)(my_lwe & 0xffu);                                  // lz.rs:136
                it.pad = 0;
                it.seq = seq;
                uint32_t mlen = 0, bump = kNoBump;
                if (my_kind == 2) {
                    uint32_t sym, aux;
                    match_fields(my_head, fm_idx, fm_len, node_lenmin(fm_pl), fm_exp, sym, aux);
                    it.symbol = (uint16_t)(sym & 0xffffu);
                    it.robits = (uint16_t)(sym >> 16);
                    it.robitlen = (uint8_t)(aux & 0xffu);
                    it.eml = (uint8_t)((aux >> 8) & 0xffu);
                    it.flags = (uint8_t)(my_al | 2u);
                    mlen = fm_len;
                    // matcher.rs:66-71.  The reference bumps the matched node and THEN writes its new node; a bump whose target
                    // is overwritten by this step or by a later step of this window (slots my_head+1 .. my_head+n_after of the
                    // bucket) is dead, and issuing it from this lane would race with that lane's record store.
                    const uint32_t n_after = (uint32_t)__popc(same_c & ~lt_mask);        // steps of this window in bucket c from this one on
                    const uint32_t ahead = (fm_idx + kRing - my_head) % kRing;
                    if (node_lenmin(fm_pl) <= fm_len && !(ahead >= 1 && ahead <= n_after)) {
                        const uint32_t lm = fm_len + 1 < 127u ? fm_len + 1 : 127u;
                        nodes[c * kNodesStride + fm_idx].pl = node_pos(fm_pl) | lm << 25;
                        bump = fm_idx;
                    }
                } else {
                    it.robits = 0;
                    it.robitlen = 0;
                    it.eml = 0;
                    it.flags = (uint8_t)my_al;
                    it.symbol = my_kind == 1 ? (uint16_t)kWordSym : (uint16_t)(cur2 & 0xffu);
                }
                items[n_items + (uint32_t)__popc(stepmask & lt_mask)] = it;
                uint4 r0 = lds_volatile_v4(&P->rec[0]);
                const uint4 r1 = lds_volatile_v4(&P->rec[4]), r2 = lds_volatile_v4(&P->rec[8]), r3 = lds_volatile_v4(&P->rec[12]);
                r0.y = (r0.y & 0xffffff00u) | mlen;                                      // expected = match_len (matcher.rs:75)
                uint4 *dst = reinterpret_cast<uint4 *>(nodes + c * kNodesStride + new_head);
                dst[0] = r0;
                dst[1] = r1;
                dst[2] = r2;
                dst[3] = r3;
                // (no fence between the record and the head: a builder that sees the new head early meets an insert of its
                // own hash entry that is younger than its nb0, which the validation rejects anyway)
                heads[c * kHeadsStride + e] = (int16_t)new_head;
                sh.blog[c][(head0 + (uint32_t)__popc(same_c & lt_mask)) % kBlog5] = e | bump << 16;
                if (!(same_ctx & lt_mask)) sh.cnt[ctx] = cnt_base + (uint32_t)__popc(same_ctx);
            }
            fence5<CL>(); // THE release: every ring store of this window is visible before the insert counts say so
            __syncwarp();
            if (is_step && !(same_c & lt_mask)) v_bcnt[c] = head0 + (uint32_t)__popc(same_c);
            n_items += n_new;
            spos += cur;
            if (lane == 0) *v_spos = spos; // a hint for the builders (and the licence to reuse mailbox slots of passed positions)
            // probes ahead that this window's inserts made stale for certain go back to their builders now, not one window later
            if (stale_ce && ready && !is_step && p >= spos) atomicCAS(&P->q, a.x, kNone);
            __syncwarp();
        }
        if (lane == 0) {
            *v_done = 1;
            if (wd_code) {
                S->status = -7;
                S->dbg[6] = (unsigned long long)wd_code;
                S->dbg[7] = spos;
            }
            S->spos = spos;
            S->after_literal = after_literal;
            S->n_items = n_items;
            if (STATS) {
                S->dbg[0] += n_items;
                S->dbg[1] += n_windows;
                S->dbg[2] += n_empty;
                S->dbg[3] += n_inval;
            }
        }
    } else if (rank > 0 || !PARK || (warp & 3) != 0) {
        // ------------------------------------------------------------------ builders (lock-free: see parse_v4.cuh)
        constexpr bool CG = CL; // uniform code for local and remote builders of a cluster: all of them read the rings L2-only
        const int bi = rank == 0 ? (PARK ? warp - 1 - (warp >> 2) : warp - 1) : NB_LOCAL + (rank - 1) * NW + warp;
        unsigned long long nbuild = 0, nabort = 0;
        uint32_t gen = 0;
        for (;;) {
            if (bcast0(ld_vol_u32(v_done))) break;
            const uint32_t sp = bcast0(ld_vol_u32(v_spos));
            const uint32_t q1 = sp + ((uint32_t)bi + NB - sp % NB) % NB; // nearest position >= sp owned by this builder
            uint32_t q = kNone;
#pragma unroll
            for (int t = 0; t < SL; t++) {
                const uint32_t qt = q1 + (uint32_t)t * NB;
                if (q != kNone || qt >= sbuf_len || qt >= sp + kLookahead5) continue;
                const uint32_t have = bcast0(ld_vol_u32(&sh.pub[(qt % NB) * SL + (qt / NB) % SL].q));
                if (have == kNone || (have & kPosMask) != qt) q = qt;
            }
            if (q == kNone) {
                // Nothing to build.  Before sleeping, look whether one of this builder's published probes went stale (the same
                // test the resolver makes when the probe enters its window): a probe invalidated by an insert that is still
                // tens of steps away from it can be rebuilt before the parse arrives, instead of costing it one build latency.
                bool stale = false;
#pragma unroll
                for (int t = 0; t < SL; t++) {
                    const uint32_t qt = q1 + (uint32_t)t * NB;
                    if (qt >= sbuf_len || qt >= sp + kLookahead5) continue;
                    Probe4 &pb = sh.pub[(qt % NB) * SL + (qt / NB) % SL];
                    volatile Probe4 *vpb = &pb;
                    const uint32_t w0 = bcast0(ld_vol_u32(&vpb->q));
                    const uint32_t nb_p = bcast0(ld_vol_u32(&vpb->nb0)), ces = bcast0(ld_vol_u32(&vpb->ces)), fmw = bcast0(ld_vol_u32(&vpb->fm));
                    const uint32_t pc = ces & 0xffu, pe = (ces >> 8) & 0x1fffu, psafe = ces >> 21;
                    const uint32_t dnew = bcast0(ld_vol_u32(&v_bcnt[pc])) - nb_p;
                    bool bad = dnew > (uint32_t)kBlog5 || dnew > psafe;
                    if (!bad && (uint32_t)lane < dnew) {
                        const uint32_t ent = ld_vol_u32(&sh.blog[pc][(nb_p + (uint32_t)lane) % kBlog5]);
                        bad = (ent & 0xffffu) == pe || ((fmw & 0xffu) > 0 && (ent >> 16) == ((fmw >> 8) & 0xfffu));
                    }
                    if (__any_sync(kFull, bad && (w0 & kPosMask) == qt && w0 != kNone)) {
                        if (lane == 0) vpb->q = kNone; // (the resolver's compare-and-swap writes the same value)
                        stale = true;
                        if (STATS) nabort += 1ull << 32;
                    }
                }
                if (!stale) __nanosleep(128); // leave the issue slots to the others
                continue;
            }
            Probe4 &pub = sh.pub[(q % NB) * SL + (q / NB) % SL];
            volatile Probe4 *vpub = &pub;
            if (STATS) nbuild++;
            if (lane == 0) vpub->q = kNone; // (ordered before the probe's fields by the fence that follows the bcnt read)
            const uint32_t prev4 = ld_u32_unaligned(win, q - 4); // bytes q-4 .. q-1
            const uint32_t cur4 = ld_u32_unaligned(win, q);      // bytes q .. q+3
            ORZ_COST(120); // window bytes (L1) + hashing
            const uint32_t b_m1 = prev4 >> 24, b_m2 = (prev4 >> 16) & 0xffu, b_m3 = (prev4 >> 8) & 0xffu, b_m4 = prev4 & 0xffu;
            const uint32_t c = (b_m1 & 0x7fu) | is_alnum(b_m2) << 7;                              // hash1(q-1)  lz.rs:481-486
            const uint32_t h2 = (b_m1 & 0x7fu) | ((b_m2 & 0x7fu) | is_alnum(b_m3) << 7) << 7;     // hash2(q-1)  lz.rs:488-492
            const uint32_t e0 = hash_dword(cur4) % kHashSize;
            const uint32_t nb0 = bcast0(ld_vol_u32(&v_bcnt[c]));
            fence5<CL>();
            const int h0 = (int)bcast0((uint32_t)(int)ld_head<CG>(heads + c * kHeadsStride + e0));
            ORZ_COST(450); // heads entry: one L2 round trip
            ProbeResult R;
            int f_next;
            uint32_t safe;
            if (!probe_chain5<CG>(win, nodes + c * kNodesStride, h0, q, sbuf_len, depth0, depth1, depth2, lane, v_spos, nb0 % kRing, R, f_next,
                              safe)) {
                if (STATS) nabort++;
                continue;
            }
            // words updates of the two possible outcomes (lz.rs:208,233): words[hash2(s'-3)] = bytes s'-2, s'-1
            const uint32_t hw_lit = (b_m2 & 0x7fu) | ((b_m3 & 0x7fu) | is_alnum(b_m4) << 7) << 7; // s' = q+1
            const uint32_t wv_lit = b_m1 | (cur4 & 0xffu) << 8;
            uint32_t hw_m = 0, wv_m = 0;
            if (R.fm_len > 0) { // s' = q + len
                const uint32_t t4 = ld_u32_unaligned(win, q + R.fm_len - 5); // bytes s'-5 .. s'-2
                const uint32_t t2 = ld_u32_unaligned(win, q + R.fm_len - 2); // bytes s'-2 .. s'+1
                const uint32_t x3 = (t4 >> 16) & 0xffu, x4 = (t4 >> 8) & 0xffu, x5 = t4 & 0xffu;
                hw_m = (x3 & 0x7fu) | ((x4 & 0x7fu) | is_alnum(x5) << 7) << 7;
                wv_m = t2 & 0xffffu;
            }
            ORZ_COST(200); // resolution + publish
            {
                // ring-slot image as 32 int16 (as v3): [0,1] pl = q, [2] expected (patched at commit), [3] next[0] = h0,
                // [4..31] next[1..28] = the old head's next[0..27]; lane L>=1 holds the old head's next[L-1] in f_next
                const int up = __shfl_sync(kFull, f_next, lane >= 3 ? lane - 3 : 1);
                int hv;
                if (lane == 0) hv = (int)(q & 0xffffu);
                else if (lane == 1) hv = (int)((q & 0x1ffffffu) >> 16);
                else if (lane == 2) hv = 0;
                else if (lane == 3) hv = h0;
                else hv = up;
                reinterpret_cast<int16_t *>(pub.rec)[lane] = (int16_t)hv;
            }
            if (lane == 0) {
                const uint32_t m1 = R.max_d1 < 0 ? 0u : (R.max_d1 > 255 ? 255u : (uint32_t)R.max_d1);
                const uint32_t m2 = R.max_d2 < 0 ? 0u : (R.max_d2 > 255 ? 255u : (uint32_t)R.max_d2);
                pub.nb0 = nb0;
                pub.ces = c | e0 << 8 | (safe > 63u ? 63u : safe) << 21;
                pub.fm = (uint32_t)R.fm_len | (uint32_t)R.fm_idx << 8 | (R.fm_exp & 0xffu) << 20;
                pub.mx = m1 | m2 << 8;
                pub.hc = h2 | (cur4 & 0xffffu) << 16;
                pub.w_lit = hw_lit | wv_lit << 16;
                pub.w_m = hw_m | wv_m << 16;
                pub.fm_pl = R.fm_pl;
            }
            fence5<CL>();
            __syncwarp();
            gen = (gen + 1) & 0x3fu;
            if (lane == 0) vpub->q = q | gen << 25;
        }
        if (STATS && lane == 0) {
            atomicAdd(&S->dbg[4], nbuild);
            atomicAdd(&S->dbg[5], nabort);
        }
    }
    __syncthreads();
    if (CL) cluster_barrier(); // no remote builder touches CTA 0's shared memory after this point
    if (rank == 0) {
        for (int i = tid; i < 32768 / 8; i += NT) reinterpret_cast<uint4 *>(S->words)[i] = reinterpret_cast<const uint4 *>(sh.words)[i];
        for (int i = tid; i < kNumCtx; i += NT) S->ctx_count[i] = sh.cnt[i];
        for (int i = tid; i < kNumBuckets; i += NT) S->bhead[i] = sh.bcnt[i] % kRing;
    }
}

} // namespace orz
