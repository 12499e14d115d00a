// parse_v6.cuh — stage A: parse_v5.cuh with SUB-WARP builders (opt-in, ORZ_B200_PARSE=6; included by encode_kernels.cu).
//
// v4 on a B200 is bound by builder throughput: every position needs a probe (4.3 builds per step), a build occupies a
// whole warp for ~4500 cycles, and most of its 32 lanes idle — find_match visits 8 chain nodes on average (enwik8, -l2)
// while a gather round offers 30.  Here a builder warp is split into NG = 32/G groups of G lanes (G = 8 or 16; 32
// reproduces v5) and every group builds its own probe, so one warp instruction advances NG probes: fewer issue slots
// per probe (what matters with many streams per SM) and NG probes in flight per warp (what matters for one stream).
// A chain is walked G nodes at a time: the group keeps the 64-byte record of the current snapshot source in
// registers (16/G words per lane) and takes its 29 links segment by segment, so a further segment costs the same two
// round trips (candidate records, candidate bytes) as a v4 round and no extra load for the links.  All collectives stay
// full-mask with width G (groups never diverge: loops run while ANY group needs them, finished groups are predicated
// off).  The resolver is v5's, unchanged; mailbox owners are groups instead of warps.
// Written and checked in the emulator in round 1 (no GPU minutes were left); round 2 measures it.
#pragma once

namespace orz {

template <int G>
struct Parse6Words {
    static constexpr int WPL = G >= 16 ? 1 : 16 / G; // record words held per lane
};

template <int G>
__device__ __forceinline__ uint32_t gshfl(uint32_t v, int src) { return __shfl_sync(kFull, v, src, G); }
template <int G>
__device__ __forceinline__ unsigned gballot(bool p, int g0) {
    const unsigned b = __ballot_sync(kFull, p) >> g0;
    return G == 32 ? b : (b & ((1u << (G & 31)) - 1u));
}
template <int G>
__device__ __forceinline__ int gmax(int v) {
#pragma unroll
    for (int o = G / 2; o; o >>= 1) {
        const int other = __shfl_xor_sync(kFull, v, o, G);
        v = other > v ? other : v;
    }
    return v;
}
// word k (0..15) of the 64-byte record that the group holds in rw[] (lane sl holds words sl, sl+G, ...); k < 0 -> 0
template <int G>
__device__ __forceinline__ uint32_t fetch_word(const uint32_t *rw, int k) {
    const int kk = k < 0 ? 0 : k;
    uint32_t v = gshfl<G>(rw[0], kk % G);
#pragma unroll
    for (int s = 1; s < Parse6Words<G>::WPL; s++) {
        const uint32_t vs = gshfl<G>(rw[s], kk % G);
        if (kk / G == s) v = vs;
    }
    return k < 0 ? 0u : v;
}

// One probe per group: matcher.rs:135-192 (find_match at q, depth0) and matcher.rs:194-228 as the lazy maxima, the
// `safe` distance of parse_v4.cuh, and rw0 = the words of the chain head's record (for the ring-slot image).
// Called by the whole warp; `act` says whether this lane's group has a probe to build.  Returns false for a group
// whose walk was abandoned because the parse already moved past q.
template <int G, bool CG>
__device__ __forceinline__ bool probe_group(const uint8_t *win, const ENode *bn, int h0, uint32_t q, uint32_t buf_len, int depth0, int depth1,
                                            int depth2, int lane, int sl, int g0, const volatile uint32_t *s_spos, uint32_t head, bool act,
                                            ProbeResult &R, uint32_t &safe, uint32_t *rw0) {
    constexpr int WPL = Parse6Words<G>::WPL;
    R.fm_len = 0;
    R.fm_idx = 0;
    R.fm_pl = 0;
    R.fm_exp = 0;
    R.max_d1 = -1;
    R.max_d2 = -1;
    safe = 0xffffu;
    uint32_t rw[WPL];
#pragma unroll
    for (int s = 0; s < WPL; s++) rw[s] = rw0[s] = 0xffffffffu; // no chain: every link of the image is -1
    bool walking = act && h0 >= 0;
    int x_idx = h0;
    if (walking) {
        const uint32_t *r = reinterpret_cast<const uint32_t *>(bn + h0);
#pragma unroll
        for (int s = 0; s < WPL; s++)
            if (sl + s * G < 16) rw[s] = rw0[s] = ld_ring<CG>(r + sl + s * G);
    }
    ORZ_COST(450);
    uint32_t x_pl = fetch_word<G>(rw, 0), x_exp = fetch_word<G>(rw, 1) & 0xffu;
    int dmax = depth0 > depth1 ? depth0 : depth1;
    dmax = dmax > depth2 ? dmax : depth2;
    if (dmax <= 0) walking = false;
    uint32_t my_dist = 0xffffu;
    if (act && h0 >= 0 && sl == 0) {
        const uint32_t d = ((uint32_t)h0 + kRing - head) % kRing;
        my_dist = d == 0 ? (uint32_t)kRing : d;
    }
    int max_len = kMatchMin - 1;
    uint32_t exp_best = kMatchMin, pl_best = (uint32_t)kMatchMin << 25;
    int idx_best = 0;
    int seen = 0;        // chain nodes consumed by earlier segments
    int soff = 0;        // links of the current source record consumed
    uint32_t tail_pos = 0;
    bool fm_done = depth0 <= 0;
    bool first = true;
    bool aborted = false;

    while (__any_sync(kFull, walking)) {
        const uint32_t sp_now = gshfl<G>(ld_vol_u32(s_spos), 0); // sub-lane 0's reading for the whole group
        if (walking && sp_now > q) {
            walking = false;
            aborted = true;
        }
        // ---- this lane's chain element: the source record itself (first segment, sub-lane 0) or one of its links
        const int li = soff + sl - (first ? 1 : 0);
        const int cap_nodes = dmax - seen < G ? dmax - seen : G;
        const bool want = walking && sl < cap_nodes && li < kSnap;
        const int j16 = (li < 0 ? 0 : li) + 3; // int16 index of next[li] in the record
        const uint32_t lw = fetch_word<G>(rw, j16 >> 1);
        const int link = (int)(int16_t)((j16 & 1) ? (lw >> 16) : (lw & 0xffffu));
        const bool head_lane = walking && first && sl == 0;
        bool ok = false;
        uint32_t my_pl = 0, my_exp = 0;
        int my_idx = -1;
        if (head_lane) {
            my_pl = x_pl;
            my_exp = x_exp;
            my_idx = x_idx;
            ok = true;
        } else if (want && li >= 0 && link >= 0) {
            const uint32_t *rr = reinterpret_cast<const uint32_t *>(bn + link);
            my_pl = ld_ring<CG>(rr);
            my_exp = ld_ring<CG>(rr + 1) & 0xffu;
            my_idx = link;
            ok = true;
        }
        ORZ_COST(450);
        // matcher.rs:171-179: link != -1, positions strictly decrease
        const uint32_t up_pos = __shfl_up_sync(kFull, node_pos(my_pl), 1, G);
        const uint32_t prev_pos = sl == 0 ? tail_pos : up_pos;
        const bool link_ok = head_lane || (ok && node_pos(my_pl) < prev_pos);
        const unsigned gmask = G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u);
        const unsigned bad = ~gballot<G>(link_ok, g0) & gmask;
        const int nv = bad ? (__ffs(bad) - 1) : G;       // leading valid elements of this segment
        if (walking && sl <= nv && my_idx >= 0) {        // the read set: every record that decided something
            const uint32_t d = ((uint32_t)my_idx + kRing - head) % kRing;
            const uint32_t dd = d == 0 ? (uint32_t)kRing : d;
            my_dist = dd < my_dist ? dd : my_dist;
        }
        const int gi = seen + sl;
        const bool have = walking && sl < nv;
        int lcp = -1;
        if (have) lcp = lcp_win_win(win, node_pos(my_pl), q);
        ORZ_COST(500);
        {
            const int m1 = gmax<G>((have && gi < depth1) ? lcp : -1), m2 = gmax<G>((have && gi < depth2) ? lcp : -1);
            if (walking) {
                R.max_d1 = max(R.max_d1, m1);
                R.max_d2 = max(R.max_d2, m2);
            }
        }
        // ---- sequential semantics of matcher.rs:149-181 over this segment's candidates (as v3, per group)
        {
            const bool valid = have && gi < depth0;
            bool looping = walking && !fm_done;
            int cur = -1;
            while (__any_sync(kFull, looping)) {
                const unsigned after = (cur >= 31) ? 0u : (0xffffffffu << (cur + 1));
                const unsigned upd = gballot<G>(looping && valid && lcp > max_len, g0) & after;
                const int u = upd ? (__ffs(upd) - 1) : 32;
                unsigned brk = gballot<G>(looping && valid && exp_best > 0 && lcp < max_len - 3 && lcp > (int)exp_best, g0) & after;
                if (u < 32) brk &= (1u << u) - 1u;
                bool eq = false;
                if (looping && ((brk >> sl) & 1u))
                    eq = ld_u32_unaligned(win, node_pos(my_pl) + max_len - 3) == ld_u32_unaligned(win, q + max_len - 3);
                const bool any_eq = gballot<G>(eq, g0) != 0;
                const int us = u < G ? u : 0;
                const int n_len = (int)gshfl<G>((uint32_t)lcp, us);
                const uint32_t n_exp = gshfl<G>(my_exp, us), n_pl = gshfl<G>(my_pl, us);
                const int n_idx = (int)gshfl<G>((uint32_t)my_idx, us);
                if (looping) {
                    if (any_eq) {
                        fm_done = true;
                        looping = false;
                    } else if (u >= G) {
                        looping = false;
                    } else {
                        max_len = n_len;
                        exp_best = n_exp;
                        pl_best = n_pl;
                        idx_best = n_idx;
                        if (max_len == kMatchMax || (exp_best > 0 && max_len > (int)exp_best)) { // matcher.rs:162-167
                            fm_done = true;
                            looping = false;
                        }
                        cur = u;
                    }
                }
            }
        }
        // ---- continue, switch the source record, or stop
        const int n_wanted = min(cap_nodes, kSnap - soff + (first ? 1 : 0));
        const int last = nv > 0 ? nv - 1 : 0;
        const uint32_t last_pos = gshfl<G>(node_pos(my_pl), last);
        const int last_idx = (int)gshfl<G>((uint32_t)my_idx, last);
        const uint32_t last_pl = gshfl<G>(my_pl, last), last_exp = gshfl<G>(my_exp, last);
        bool reload = false;
        if (walking) {
            seen += nv;
            if (seen >= depth0) fm_done = true;
            const bool lazy_done = (seen >= depth1 || R.max_d1 >= kMatchMax / 2 + 2) && (seen >= depth2 || R.max_d2 >= kMatchMax / 2 + 2);
            if (nv < n_wanted || (fm_done && lazy_done) || seen >= dmax) {
                walking = false; // the chain ended inside this segment, or nothing more is needed
            } else {
                tail_pos = last_pos;
                soff += nv - (first ? 1 : 0);
                first = false;
                if (soff >= kSnap) { // the snapshot is used up: its last node becomes the source (it is not a candidate again)
                    x_idx = last_idx;
                    x_pl = last_pl;
                    x_exp = last_exp;
                    soff = 0;
                    reload = true;
                }
            }
        }
        if (__any_sync(kFull, reload)) {
            if (reload) {
                const uint32_t *r = reinterpret_cast<const uint32_t *>(bn + x_idx);
#pragma unroll
                for (int s = 0; s < WPL; s++)
                    if (sl + s * G < 16) rw[s] = ld_ring<CG>(r + sl + s * G);
            }
            ORZ_COST(450);
        }
    }
    if (max_len >= kMatchMin && q + max_len < buf_len) { // matcher.rs:183
        R.fm_len = max_len;
        R.fm_idx = idx_best;
        R.fm_pl = pl_best;
        R.fm_exp = exp_best;
    }
#pragma unroll
    for (int o = G / 2; o; o >>= 1) {
        const uint32_t other = __shfl_xor_sync(kFull, my_dist, o, G);
        my_dist = other < my_dist ? other : my_dist;
    }
    safe = my_dist - 1;
    return !aborted;
}

// As orz_parse5_kernel, with G-lane builder groups.  K CTAs of NW warps per stream (a cluster when K > 1).  CTA 0: warp 0 is the resolver, the other warps build
// (PARK: except those on the resolver's SM sub-partition); CTAs 1..K-1: every warp builds.  SL mailbox slots per
// builder (look-ahead = NB * SL positions).
template <int NW, int SL, int K, bool PARK, int G>
struct Parse6Dims {
    static constexpr int NG = 32 / G;                                          // probes in flight per builder warp
    static constexpr int NBW_LOCAL = PARK ? NW - 1 - (NW - 1) / 4 : NW - 1;    // builder warps in CTA 0
    static constexpr int NB = (NBW_LOCAL + (K - 1) * NW) * NG;                 // builder GROUPS = mailbox owners
};

template <int NW, int SL, int K, bool PARK, int G, bool STATS>
__global__ void __launch_bounds__(NW * 32) orz_parse6_kernel(EncStream *streams) {
    constexpr int NG = Parse6Dims<NW, SL, K, PARK, G>::NG;
    constexpr int NBW_LOCAL = Parse6Dims<NW, SL, K, PARK, G>::NBW_LOCAL;
    constexpr int NB = Parse6Dims<NW, SL, K, PARK, G>::NB;
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
        // ------------------------------------------------------------------ builders: NG groups of G lanes per warp, one probe each
        constexpr bool CG = CL;
        const int bw = rank == 0 ? (PARK ? warp - 1 - (warp >> 2) : warp - 1) : NBW_LOCAL + (rank - 1) * NW + warp;
        const int sl = lane % G, g0 = lane - sl;
        const int gb = bw * NG + lane / G; // this group's index among all builder groups of the stream
        unsigned long long nbuild = 0, nabort = 0;
        uint32_t gen = 0;
        for (;;) {
            if (bcast0(ld_vol_u32(v_done))) break;
            const uint32_t sp = bcast0(ld_vol_u32(v_spos));
            const uint32_t q1 = sp + ((uint32_t)gb + NB - sp % NB) % NB; // nearest position >= sp owned by this group
            uint32_t q = kNone;
#pragma unroll
            for (int t = 0; t < SL; t++) {
                const uint32_t qt = q1 + (uint32_t)t * NB;
                const uint32_t qs = qt < sbuf_len ? qt : q1; // keep the address in range; the result is ignored then
                const uint32_t have = gshfl<G>(ld_vol_u32(&sh.pub[(qs % NB) * SL + (qs / NB) % SL].q), 0);
                if (q == kNone && qt < sbuf_len && qt < sp + kLookahead5 && (have == kNone || (have & kPosMask) != qt)) q = qt;
            }
            bool act = q != kNone;
            if (!__any_sync(kFull, act)) {
                // Nothing to build in this warp.  Before sleeping every group re-checks its published probes with the resolver's
                // own staleness test (parse_v5.cuh): a probe invalidated by an insert that is still tens of steps away from it
                // can be rebuilt before the parse arrives.
                bool stale = false;
#pragma unroll
                for (int t = 0; t < SL; t++) {
                    const uint32_t qt = q1 + (uint32_t)t * NB;
                    const bool in_range = qt < sbuf_len && qt < sp + kLookahead5;
                    const uint32_t qs = in_range ? qt : q1 % NB; // (any valid slot address; the result is ignored)
                    Probe4 &pb = sh.pub[(qs % NB) * SL + (qs / NB) % SL];
                    volatile Probe4 *vpb = &pb;
                    const uint32_t w0 = gshfl<G>(ld_vol_u32(&vpb->q), 0);
                    const uint32_t nb_p = gshfl<G>(ld_vol_u32(&vpb->nb0), 0), ces = gshfl<G>(ld_vol_u32(&vpb->ces), 0);
                    const uint32_t fmw = gshfl<G>(ld_vol_u32(&vpb->fm), 0);
                    const bool mine = in_range && w0 != kNone && (w0 & kPosMask) == qt;
                    const uint32_t pc = ces & 0xffu, pe = (ces >> 8) & 0x1fffu, psafe = ces >> 21;
                    const uint32_t dnew = gshfl<G>(ld_vol_u32(&v_bcnt[pc]), 0) - nb_p;
                    bool bad = dnew > (uint32_t)kBlog5 || dnew > psafe;
                    if (mine && !bad)
                        for (uint32_t k = (uint32_t)sl; k < dnew; k += G) {
                            const uint32_t ent = ld_vol_u32(&sh.blog[pc][(nb_p + k) % kBlog5]);
                            if ((ent & 0xffffu) == pe || ((fmw & 0xffu) > 0 && (ent >> 16) == ((fmw >> 8) & 0xfffu))) bad = true;
                        }
                    const bool gbad = gballot<G>(mine && bad, g0) != 0;
                    if (gbad) {
                        if (sl == 0) vpb->q = kNone; // (the resolver's compare-and-swap writes the same value)
                        stale = true;
                    }
                }
                if (!__any_sync(kFull, stale)) __nanosleep(128); // leave the issue slots to the others
                continue;
            }
            const uint32_t qq = act ? q : sp; // inactive groups run along on a harmless position, publish nothing
            Probe4 &pub = sh.pub[(qq % NB) * SL + (qq / NB) % SL];
            volatile Probe4 *vpub = &pub;
            if (STATS && act && sl == 0) nbuild++;
            if (act && sl == 0) vpub->q = kNone; // (ordered before the probe's fields by the fence that follows the bcnt read)
            const uint32_t prev4 = ld_u32_unaligned(win, qq - 4); // bytes q-4 .. q-1
            const uint32_t cur4 = ld_u32_unaligned(win, qq);      // bytes q .. q+3
            ORZ_COST(120);
            const uint32_t b_m1 = prev4 >> 24, b_m2 = (prev4 >> 16) & 0xffu, b_m3 = (prev4 >> 8) & 0xffu, b_m4 = prev4 & 0xffu;
            const uint32_t c = (b_m1 & 0x7fu) | is_alnum(b_m2) << 7;                              // hash1(q-1)  lz.rs:481-486
            const uint32_t h2 = (b_m1 & 0x7fu) | ((b_m2 & 0x7fu) | is_alnum(b_m3) << 7) << 7;     // hash2(q-1)  lz.rs:488-492
            const uint32_t e0 = hash_dword(cur4) % kHashSize;
            const uint32_t nb0 = gshfl<G>(ld_vol_u32(&v_bcnt[c]), 0);
            fence5<CL>();
            const uint32_t hv = act ? (uint32_t)ld_head<CG>(heads + c * kHeadsStride + e0) : 0xffffffffu;
            const int h0 = (int)gshfl<G>(hv, 0); // (one shuffle for all groups: collectives never sit under a group-dependent branch)
            ORZ_COST(450);
            ProbeResult R;
            uint32_t safe, rw0[Parse6Words<G>::WPL];
            const bool done_ok = probe_group<G, CG>(win, nodes + c * kNodesStride, h0, qq, sbuf_len, depth0, depth1, depth2, lane, sl, g0,
                                                    v_spos, nb0 % kRing, act, R, safe, rw0);
            if (STATS && act && !done_ok && sl == 0) nabort++;
            act = act && done_ok;
            // words updates of the two possible outcomes (lz.rs:208,233): words[hash2(s'-3)] = bytes s'-2, s'-1
            const uint32_t hw_lit = (b_m2 & 0x7fu) | ((b_m3 & 0x7fu) | is_alnum(b_m4) << 7) << 7; // s' = q+1
            const uint32_t wv_lit = b_m1 | (cur4 & 0xffu) << 8;
            uint32_t hw_m = 0, wv_m = 0;
            if (R.fm_len > 0) { // s' = q + len
                const uint32_t t4 = ld_u32_unaligned(win, qq + R.fm_len - 5); // bytes s'-5 .. s'-2
                const uint32_t t2 = ld_u32_unaligned(win, qq + R.fm_len - 2); // bytes s'-2 .. s'+1
                const uint32_t x3 = (t4 >> 16) & 0xffu, x4 = (t4 >> 8) & 0xffu, x5 = t4 & 0xffu;
                hw_m = (x3 & 0x7fu) | ((x4 & 0x7fu) | is_alnum(x5) << 7) << 7;
                wv_m = t2 & 0xffffu;
            }
            ORZ_COST(200);
            // ring-slot image (as v3): int16 [0,1] pl = q, [2] expected (patched at commit), [3] next[0] = h0, [4..31] = the old
            // head's next[0..27], i.e. the old record's int16 [3..30] moved up by one.  The group holds the old record's words.
#pragma unroll
            for (int s = 0; s < Parse6Words<G>::WPL; s++) {
                const int kw = sl + s * G; // word of the image this lane writes (G = 32: lanes 16.. write nothing)
                const uint32_t prevw = fetch_word<G>(rw0, kw - 1);
                uint32_t nw;
                if (kw == 0) nw = qq & kPosMask;
                else if (kw == 1) nw = ((uint32_t)h0 & 0xffffu) << 16;
                else nw = (prevw >> 16) | (rw0[s] << 16);
                if (act && kw < 16) pub.rec[kw] = nw;
            }
            if (act && sl == 0) {
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
            if (act && sl == 0) vpub->q = qq | gen << 25;
        }
        if (STATS && sl == 0) {
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
