// parse_v7.cuh — stage A, round-2 formulation (included by encode_kernels.cu; the default for few streams).
//
// lz.rs:131-235 is one serial chain per stream: which positions become ring nodes depends on every earlier decision.
// v3..v6 (round 1) kept the reference's ring + hash-chain layout in HBM, walked it from builder warps and let the warp
// that owns the chain also write every ring record, head, item and log entry — 1180 SM cycles per step on a B200
// against 490 for one host core.  This formulation changes the data layout and the split of work:
//
//  * PER-KEY CANDIDATE LISTS instead of hash chains.  The candidates of find_match / has_lazy_match at position y are
//    the ring nodes of bucket c = hash1(y-1) whose hash entry is e = hash_dword(y) % 4627, newest first
//    (matcher.rs:135-228).  In the reference they are reached by one dependent load per hop (heads -> nexts -> nexts);
//    a chain hop that lands on a re-used ring slot stops the walk ("positions strictly decrease", matcher.rs:176-179) or,
//    for a stale `heads` entry, walks nodes of other hash entries whose first dword differs from y's and which therefore can
//    never be taken nor trigger an early exit.  So the walk's RESULT is exactly: the nodes of key (c,e) that are still
//    among the bucket's last 4094 inserts and whose position survived `forward`, newest first, the first `depth` of them.
//    Here every key owns a 64-entry circular list in HBM (position, expected length, insert ordinal of the bucket); a
//    node is alive iff  bucket_inserts_now - 1 - ordinal <= 4093.  A builder fetches all candidates of a position with
//    ONE gather, not a chain of dependent loads, and nothing has to be re-linked when a slot is re-used.
//    (match_len_min, the only mutable per-node field, lives in a byte array indexed by ring slot; it never influences the
//    path, only the emitted length code, so only the committer touches it.)
//
//  * A CLUSTER per stream.  CTA 0 holds, in shared memory, a ring of PROBE RECORDS for the next 512 positions, `words`,
//    the per-bucket insert counts and a log of recent steps.  All other CTAs of the cluster are BUILDERS: warp b builds
//    the record of every position y = b (mod NBW) ahead of the parse — static fields, candidate gather, LCPs, the
//    ballot-resolved semantics of matcher.rs:135-192 and the two lazy maxima — against the state committed up to a
//    frontier F, and stores it into CTA 0's ring through DSMEM.
//
//  * The RESOLVER (warp 0 of CTA 0) does nothing but decide.  Lane l owns position spos+l.  A record built at
//    frontier F is exact at the moment its position is decided iff no step in [F, now) has the same key (a 4096-entry
//    "last step with this key hash" table answers that with one shared load) and its oldest candidate is still alive.
//    The walk through the window costs one shuffle round per step; every lane keeps its decision inputs current by
//    folding the broadcast step (bucket insert count -> robitlen<8, words update -> last_word_matched, same key -> stale).
//    Decided steps go into a shared-memory queue; a stale record is repaired in place (REPAIR: the committed state is
//    drained and the record rebuilt at frontier = now).
//
//  * The COMMITTER (warp 1 of CTA 0) drains the queue 32 steps at a time: list appends, lenmin bumps with the
//    reference's in-order semantics (matcher.rs:62-80), items (lz.rs:173-236), and publishes the new frontier.
//    None of its HBM latency is on the chain.
//
// All citations are file:line into richox/orz @29414c6.
#pragma once
#ifndef ORZ_SIMT_EMU
#include <cooperative_groups.h>
#endif

#ifdef ORZ_SIMT_EMU
#include <cstddef>
#include <cstdio>
#define V7_TRACE(...)                                              \
    do {                                                           \
        if (getenv("ORZ_V7_TRACE")) fprintf(stderr, __VA_ARGS__);  \
    } while (0)
#else
#define V7_TRACE(...) ((void)0)
#endif

namespace orz {

constexpr int kListCap = 64;           // entries per key list (circular, newest at (count-1) % 64)
constexpr int kMaxDepth7 = 48;         // deepest LZCfg depth this kernel handles (levels 0-2 use 45)
constexpr int kRing7 = 480;            // probe record slots in CTA 0's shared memory (112 builders x 4, 240 x 2, 6 x 80)
#ifndef ORZ_V7_LEAD
#define ORZ_V7_LEAD 160 // (measured on 6 MB / 100 MB of enwik8 -l2: 96: 11.7 / 12.2 MB/s, 128: 12.3, 160: 12.6 / 13.3, 224: 12.4)
#endif
#ifndef ORZ_V7_MUT_TIGHT
#define ORZ_V7_MUT_TIGHT 0u // (mutation hook of the A/B builds: a wrong ring-edge check must break the 100 MB identity test)
#endif
#ifndef ORZ_V7_REFRESH
#define ORZ_V7_REFRESH 45 // a stale record at least this far ahead of the round's base is handed back to its builder
#endif
#ifndef ORZ_V7_TEAM
#define ORZ_V7_TEAM 4
#endif
#ifndef ORZ_V7_VERIFY_TABLE
#define ORZ_V7_VERIFY_TABLE 1 // verification of a super-window with constant work per step (MATCH + a table of lane masks)
#endif
#ifndef ORZ_V7_TRACE_ALWAYS
#define ORZ_V7_TRACE_ALWAYS 0 // 1: the event trace of ORZ_B200_DBG_POS is compiled into the production instantiation too (A/B builds only)
#endif
#ifndef ORZ_V7_TEST_TEAR
#define ORZ_V7_TEST_TEAR 0 // (fault injection of the emulator tests: builders publish the stamp BEFORE the words of a record, the way the
                           // hardware was seen to; readers must find out through the checksums and wait)
#endif
#ifndef ORZ_V7_MUT_NOCHECK
#define ORZ_V7_MUT_NOCHECK 0 // (mutation hook: readers ignore the record checksums — with the fault injection on, the stream must then differ)
#endif
#ifndef ORZ_V7_PROV
#define ORZ_V7_PROV 0 // 1: every handed-over step leaves two uint4 of provenance (how its record came about) in EncStream::prov (A/B builds only)
#endif
#ifndef ORZ_V7_EDGE_EXACT
#define ORZ_V7_EDGE_EXACT 1 // a robitlen < 8 that an insert could flip ends a round only if the flip would change the decision
#endif
constexpr uint32_t kLeadTarget7 = ORZ_V7_LEAD; // how far ahead of the parse records are built: every position of lead is a chance to go stale
constexpr int kExtCand = 48;           // candidates per record kept for repairs
constexpr int kSumWords = 12;          // summary words per record
constexpr int kExtWords = 2 + kExtCand; // F_ext, ord0, candidates
constexpr int kQueue7 = 512;           // step queue / recent-step log entries
constexpr int kInRing = 4096;           // bytes of the window mirrored in CTA 0's shared memory
constexpr uint32_t kInAhead = 1536;    // ... from 2048 bytes behind the parse position to this far ahead of it
constexpr uint32_t kOrdMask = 0x7fffffffu;
constexpr uint32_t kTyLit = 0, kTyWord = 1, kTyMatch = 2;
constexpr uint32_t kRefreshAhead = ORZ_V7_REFRESH;
constexpr int kTeam7 = ORZ_V7_TEAM;            // decide warps beside the resolver: 4 x 30 positions per round
constexpr int kTeamThreads = (kTeam7 + 1) * 32; // ... and the resolver: the threads of the two round barriers

// A key list entry: pos:25 | expected(low 7 bits):7  ,  ordinal:31 | expected(bit 7):1
__device__ __forceinline__ uint2 kentry_make(uint32_t pos, uint32_t exp, uint32_t ord) {
    return make_uint2((pos & 0x1ffffffu) | (exp & 0x7fu) << 25, (ord & kOrdMask) | (exp >> 7) << 31);
}
__device__ __forceinline__ uint32_t kentry_pos(uint2 e) { return e.x & 0x1ffffffu; }
__device__ __forceinline__ uint32_t kentry_exp(uint2 e) { return (e.x >> 25) | (e.y >> 31) << 7; }
__device__ __forceinline__ uint32_t kentry_ord(uint2 e) { return e.y & kOrdMask; }

#ifdef ORZ_SIMT_EMU
template <class T>
__device__ __forceinline__ T *map_to_cta0(T *p) { return reinterpret_cast<T *>(simt::cur->cluster_smem[0] + (reinterpret_cast<uint8_t *>(p) - simt::cur->dyn_smem)); }
__device__ __forceinline__ void cluster_barrier() { simt::cluster_sync(); }
__device__ __forceinline__ unsigned cluster_cta_rank() { return simt::cur->cta_rank; }
#else
template <class T>
__device__ __forceinline__ T *map_to_cta0(T *p) { return cooperative_groups::this_cluster().map_shared_rank(p, 0); }
__device__ __forceinline__ void cluster_barrier() { cooperative_groups::this_cluster().sync(); }
__device__ __forceinline__ unsigned cluster_cta_rank() { return cooperative_groups::this_cluster().block_rank(); }
#endif

// Named barriers (PTX bar.sync / bar.arrive): a warp that waits at one takes no issue slot and makes no shared-memory traffic —
// unlike a polling loop, which measurably slows a resolver warp on the same SM.
__device__ __forceinline__ void bar_sync7(int id, int nthreads) {
#ifdef ORZ_SIMT_EMU
    simt::named_barrier(id, nthreads, true);
#else
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
#endif
}
__device__ __forceinline__ void bar_arrive7(int id, int nthreads) {
#ifdef ORZ_SIMT_EMU
    simt::named_barrier(id, nthreads, false);
#else
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
#endif
}

// one 16-byte volatile load through a generic address (CTA 0's shared memory seen from any CTA of the cluster)
__device__ __forceinline__ uint4 ld_volatile_v4_generic(const void *p) {
    uint4 v;
#ifdef ORZ_SIMT_EMU
    ORZ_POLL();
    v = *reinterpret_cast<const uint4 *>(p);
#else
    asm volatile("ld.volatile.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
#endif
    return v;
}
// cluster-scope acquire / release on generic addresses (CTA 0's shared memory seen from any CTA of the cluster)
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
#ifdef ORZ_SIMT_EMU
    return *p;
#else
    uint32_t v;
    asm volatile("ld.acquire.cluster.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v) {
#ifdef ORZ_SIMT_EMU
    *p = v;
#else
    asm volatile("st.release.cluster.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}
__device__ __forceinline__ void fence_cluster7() {
#ifndef ORZ_SIMT_EMU
    asm volatile("fence.acq_rel.cluster;" ::: "memory");
#endif
}
__device__ __forceinline__ uint32_t ldcg_u32(const uint32_t *p) {
#ifdef ORZ_SIMT_EMU
    return *p;
#else
    return __ldcg(p);
#endif
}
__device__ __forceinline__ uint32_t ldcg_u8(const uint8_t *p) {
#ifdef ORZ_SIMT_EMU
    return *p;
#else
    return (uint32_t)__ldcg(p);
#endif
}

struct ParseShared7 {
    uint16_t words[32768];                  // lz.rs:52 — resolver only
    uint32_t sum[kRing7][kSumWords];        // probe summaries; position y lives in its builder's own slots: slot7<NBW>(y)
    uint32_t ext[kRing7][kExtWords];        // per record: F_ext, ord0, then lcp | expected << 8 | (ord0 - ord) << 16 per candidate
    uint4 queue[kQueue7];                   // decided steps: resolver -> committer; the tail doubles as the recent-step log
    uint32_t lastkey[4096];                 // position of the latest decided step whose key hashes here (0 = none in this launch)
    uint32_t bcnt[kNumBuckets];             // inserts into the bucket decided so far (ordinal of the next insert)
    uint32_t cnt[kNumCtx];                  // committer: items per rank-coder context in this chunk
    uint32_t inring[kInRing / 4];           // window bytes around the parse position (byte p at p % 4096), kept by the feeder warp
    uint32_t rc_pos[32], rc_exp[32], rc_ord[32]; // repair scratch: recent steps with the record's key
    // the decide team: block t covers positions team_base + 30 t + lane; for every lane the chase word and the step it would take
    uint2 pk_all[kTeam7][32];               // the chase from this lane, resolved inside the block: x = lanes stepped on, y = where it ends (see decide_block7)
    uint32_t ent_all[kTeam7 * 32][8];       // x, type|advance|lengths, key of x+1 | literal byte, key of x+2, key, words update, hash2 | words entry read, matched ordinal
    uint32_t sidx[32];                      // the super-window's steps: index into ent_all
    uint2 vkw[32];                          // ... and their (key, words update), for the verification loop
    uint32_t lcnt[kNumBuckets], icnt[kNumBuckets]; // the committers' own insert counts (they number the inserts of a batch in order)
    uint32_t vtab[512];                     // verification: lane masks of the super-window's steps by key hash [0,256) and by words index hash [256,512); all zero between rounds
    // the four words a builder polls, in one 16-byte line: one remote (DSMEM) load per poll instead of four
    uint32_t spos;                          // parse position: every step before it is decided for good (resolver -> builders)
    uint32_t lead_pos;                      // how far the resolver has looked (>= spos; a verification cut may fall back behind it)
    uint32_t fpos;                          // every step at a position < fpos is committed to HBM (committer -> builders)
    uint32_t done;
    uint32_t qhead;                         // steps queued (resolver)
    uint32_t qdone;                         // steps whose list entries are committed (list committer)
    uint32_t qitem;                         // steps whose items are written (item committer)
    uint32_t fed_hi;                        // inring holds window bytes [max(fed_lo0, fed_hi - 4096), fed_hi)
    uint32_t fed_lo0;
    uint32_t team_base;                     // first position of the decide team's round (the rounds are paced by named barriers 1 and 2)
    uint32_t torn;                          // records whose stamp was there before their words were (checksum mismatch; counted, then waited for)
    // refresh of records that went stale well ahead of the parse (bit = record slot): the decide team asks (`req`), the builder that
    // owns the slot rebuilds the record against the present frontier (`busy` while it does)
    uint32_t req[kRing7 / 32];
    uint32_t busy[kRing7 / 32];
};

static_assert(offsetof(ParseShared7, spos) % 16 == 0, "stage A v7: the polled line");

// summary words
//  0  stamp: the position the record describes | generation << 25 (a rebuild bumps it); written last, with release semantics
//  1  key = c << 13 | e   | ncand << 21 (candidates considered, <= 48) | flags << 27 (bit 0: incomplete, must be rebuilt)
//  2  hash2(y-1) | bytes(y, y+1) << 15
//  3  fm_len | lazy max over depth1 << 8 | lazy max over depth2 << 16 | expected of the matched node << 24
//  4  ordinal of the matched node
//  5  ordinal of the oldest candidate considered
//  6  F: the record is exact for every decision time T >= F as long as no step in [F, T) has its key
//  7  words update if a literal is emitted here: hash2 index | value << 15   (lz.rs:233)
//  8  words update if the match is emitted here                               (lz.rs:208)
//  9  LCP maxima without the last candidate: find_match's (depth0 - 1) | depth1 - 1 << 8 | depth2 - 1 << 16

// warp maximum of small non-negative values (LCPs; -1 = none) with one REDUX instead of five shuffles
__device__ __forceinline__ int warp_max_u(int v) { return (int)__reduce_max_sync(kFull, (unsigned)(v + 1)) - 1; }

struct FmState {
    int max_len;
    uint32_t exp_best;
    int best;     // candidate index of the best node so far
    bool done;
};

// One round (32 candidates in chain order, lane = order) of matcher.rs:149-181.  `probe_eq(max_len)` is evaluated by the
// lanes whose candidate cannot improve but whose probe dword (matcher.rs:150-153) may still match, which lets
// matcher.rs:165-167 break on them.
template <class ProbeEq>
__device__ __forceinline__ void fm_round(FmState &st, bool valid, int lcp, uint32_t exp, int cand, int lane, ProbeEq probe_eq) {
    if (st.done) return;
    int cur = -1;
    for (;;) {
        const unsigned after = (cur >= 31) ? 0u : (0xffffffffu << (cur + 1));
        const unsigned upd = __ballot_sync(kFull, valid && lcp > st.max_len) & after;
        const int u = upd ? (__ffs(upd) - 1) : 32;
        unsigned brk = __ballot_sync(kFull, valid && st.exp_best > 0 && lcp < st.max_len - 3 && lcp > (int)st.exp_best) & after;
        if (u < 32) brk &= (1u << u) - 1u;
        if (brk) {
            bool eq = false;
            if ((brk >> lane) & 1u) eq = probe_eq(st.max_len);
            if (__ballot_sync(kFull, eq)) {
                st.done = true;
                return;
            }
        }
        if (u >= 32) return;
        st.max_len = __shfl_sync(kFull, lcp, u);
        st.exp_best = __shfl_sync(kFull, exp, u);
        st.best = __shfl_sync(kFull, cand, u);
        if (st.max_len == kMatchMax || (st.exp_best > 0 && st.max_len > (int)st.exp_best)) { // matcher.rs:162-167
            st.done = true;
            return;
        }
        cur = u;
    }
}

// Checksums of a probe record (summary words 10 and 11).  The stamp alone is not a safe "record complete" signal on the hardware:
// on a B200 a reader that saw the new stamp could still read words of the slot's previous record a moment later (one 100 MB run
// in eight took a wrong decision that way, always right after a just-in-time build).  Word 10 covers summary words 1..9, word 11
// the ext words a repair reads; both are salted with the position, so a stale checksum of the previous occupant never matches.
__device__ __forceinline__ uint32_t sumcheck7(uint32_t y, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4, uint32_t w5, uint32_t w6, uint32_t w7, uint32_t w8,
                                              uint32_t w9) {
    return (w1 ^ __funnelshift_l(w2, w2, 3) ^ __funnelshift_l(w3, w3, 7) ^ __funnelshift_l(w4, w4, 11) ^ __funnelshift_l(w5, w5, 13) ^ __funnelshift_l(w6, w6, 17) ^
            __funnelshift_l(w7, w7, 19) ^ __funnelshift_l(w8, w8, 23) ^ __funnelshift_l(w9, w9, 29)) +
           y * 0x9E3779B1u + 0x5bd1e995u;
}
// the ext part: lane l holds candidate words l and 32 + l (0 where the record has none); F_ext and ord0 go in once
__device__ __forceinline__ uint32_t extcheck7(uint32_t y, uint32_t f_ext, uint32_t ord0, uint32_t c0, uint32_t c1, int lane) {
    const uint32_t mine = __funnelshift_l(c0, c0, lane & 31) ^ __funnelshift_l(c1, c1, (lane + 7) & 31);
    return __reduce_xor_sync(kFull, mine) ^ __funnelshift_l(f_ext, f_ext, 5) ^ __funnelshift_l(ord0, ord0, 9) ^ (y * 0x85EBCA6Bu + 0x2545F491u);
}

// event trace of one position (instrumented build only): see EncStream::dbgx
__device__ __forceinline__ void trace7(EncStream *S, uint32_t tag, const uint32_t *v, int n) {
    const uint32_t i = atomicAdd(&S->dbgx_n, (uint32_t)(n + 1));
    if (i + (uint32_t)n + 1u <= 1024u) {
        S->dbgx[i] = tag | (uint32_t)n << 24;
        for (int k = 0; k < n; k++) S->dbgx[i + 1 + k] = v[k];
    }
}

struct Build7 {
    EncStream *S;
    uint32_t dbg_pos;     // 0, or the window position whose records are traced (instrumented build)
    uint32_t prov_base;   // absolute input offset of window position PREMATCH (provenance log of the A/B builds)
    const uint8_t *win;
    const uint2 *klist;
    const uint32_t *kcnt;
    uint32_t *gext;       // [kRing7][kExtCand] candidate positions (HBM), for byte probes during repairs
    uint32_t sbuf_len;
    int depth0, depth1, depth2;
};

// Builds the record of position y against the state committed below frontier F.  `gc` = a lower bound of the bucket's
// insert count at the time y is decided.  Writes summary + ext through `sum`/`ext` (local or DSMEM pointers) and
// publishes the stamp last.  Warp-cooperative.
// `mode`: 0 = first build of y in this slot; 1 = the owner rebuilds the record in place while the decide team may be reading it: the stamp
// is taken away first and comes back with the next generation (a reader compares the stamp before and after its reads); 2 = rebuild
// by the resolver itself (no reader is running: same generation).
__device__ __forceinline__ void build_record7(const Build7 &B, uint32_t y, uint32_t F, const volatile uint32_t *bcnt_src, uint32_t *sum,
                                              uint32_t *ext, uint32_t slot, int lane, int mode = 0) {
    const uint8_t *win = B.win;
    const uint32_t prev4 = ld_u32_unaligned(win, y - 4); // bytes y-4 .. y-1
    const uint32_t cur4 = ld_u32_unaligned(win, y);      // bytes y .. y+3
    ORZ_COST(120);
    const uint32_t b_m1 = prev4 >> 24, b_m2 = (prev4 >> 16) & 0xffu, b_m3 = (prev4 >> 8) & 0xffu, b_m4 = prev4 & 0xffu;
    const uint32_t c = (b_m1 & 0x7fu) | is_alnum(b_m2) << 7;                          // hash1(y-1)  lz.rs:481-486
    const uint32_t h2 = (b_m1 & 0x7fu) | ((b_m2 & 0x7fu) | is_alnum(b_m3) << 7) << 7; // hash2(y-1)  lz.rs:488-492
    const uint32_t e0 = hash_dword(cur4) % kHashSize;                                  // matcher.rs:136
    const uint32_t kidx = c * kHeadsStride + e0;
    const uint32_t gc = bcast0(ld_vol_u32(&bcnt_src[c]));
    const uint32_t cnt = bcast0(ldcg_u32(&B.kcnt[kidx + (F >> 31)])); // (F < 2^25: the term is 0; it only makes the load wait for F)
    const uint32_t avail = cnt < (uint32_t)kListCap ? cnt : (uint32_t)kListCap;
    int dmax = B.depth0 > B.depth1 ? B.depth0 : B.depth1;
    dmax = dmax > B.depth2 ? dmax : B.depth2;
    const uint2 *kl = B.klist + (size_t)kidx * kListCap;
    ORZ_COST(450);

    // entries newest first: t = 0 is entry (cnt-1) % 64.  The first r of them may already belong to steps at or after F
    // (committed while this build runs): they are not part of the state "below F".
    uint2 en = make_uint2(0, 0);
    bool have = (uint32_t)lane < avail;
    if (have) en = ldcg_u2(&kl[(cnt - 1 - (uint32_t)lane) % kListCap]);
    ORZ_COST(450);
    // the entries at or above the frontier must form a prefix (newest first); anything else means the list was appended
    // to (and a slot re-used) while it was being read: the record is marked incomplete and rebuilt by a repair
    const unsigned rm = __ballot_sync(kFull, have && kentry_pos(en) >= F);
    const uint32_t r = rm == 0xffffffffu ? 32u : (uint32_t)(__ffs(~rm) - 1);
    bool incomplete = r < 32u && (rm >> r) != 0u;
    if (r > 0 && !incomplete) {
        if (r >= 32 || r + (uint32_t)dmax > (uint32_t)(kListCap - 8)) incomplete = true;
        else {
            have = r + (uint32_t)lane < avail;
            if (have) en = ldcg_u2(&kl[(cnt - 1 - r - (uint32_t)lane) % kListCap]);
            ORZ_COST(450);
            if (__ballot_sync(kFull, have && kentry_pos(en) >= F)) incomplete = true;
        }
    }
    // alive: position survived forward (pos > 0), still among the bucket's last 4094 inserts, below the frontier
    bool ok0 = !incomplete && have && lane < dmax && kentry_pos(en) != 0 && ((gc - 1 - kentry_ord(en)) & kOrdMask) <= (uint32_t)(kRing - 1);
    // candidates form a prefix (positions and ordinals decrease along the list)
    unsigned okm0 = __ballot_sync(kFull, ok0);
    int n0 = okm0 == 0xffffffffu ? 32 : (__ffs(~okm0) - 1);
    ok0 = lane < n0;
    // second round: candidates 32..47
    uint2 en1 = make_uint2(0, 0);
    bool ok1 = false;
    int n1 = 0;
    if (n0 == 32 && dmax > 32 && !incomplete) {
        const uint32_t t = r + 32u + (uint32_t)lane;
        bool have1 = t < avail && 32 + lane < dmax;
        if (have1) en1 = ldcg_u2(&kl[(cnt - 1 - t) % kListCap]);
        ORZ_COST(450);
        if (__ballot_sync(kFull, have1 && kentry_pos(en1) >= F)) incomplete = true;
        ok1 = !incomplete && have1 && kentry_pos(en1) != 0 && ((gc - 1 - kentry_ord(en1)) & kOrdMask) <= (uint32_t)(kRing - 1);
        const unsigned okm1 = __ballot_sync(kFull, ok1);
        n1 = okm1 == 0xffffffffu ? 32 : (__ffs(~okm1) - 1);
        ok1 = lane < n1;
    }
    if (incomplete) {
        ok0 = ok1 = false;
        n0 = n1 = 0;
    }
    const int ncand = n0 + n1;
    const uint32_t p0 = kentry_pos(en), p1 = kentry_pos(en1);
    int lcp0 = -1, lcp1 = -1;
    if (ok0) lcp0 = lcp_win_win(win, p0, y);
    if (ok1) lcp1 = lcp_win_win(win, p1, y);
    ORZ_COST(500);
    // lazy maxima (matcher.rs:194-228 restated: has_lazy_match(min_len) <=> longest LCP among the first `depth` nodes >= min_len)
    int m1 = warp_max_u(max((ok0 && lane < B.depth1) ? lcp0 : -1, (ok1 && 32 + lane < B.depth1) ? lcp1 : -1));
    int m2 = warp_max_u(max((ok0 && lane < B.depth2) ? lcp0 : -1, (ok1 && 32 + lane < B.depth2) ? lcp1 : -1));
    // the same maxima over one candidate fewer (and the plain LCP maximum of find_match's candidates): what a repair needs when
    // exactly one newer node of this key has appeared since the frontier (it takes the first place of the sequence)
    const int mf = warp_max_u(max((ok0 && lane < B.depth0 - 1) ? lcp0 : -1, (ok1 && 32 + lane < B.depth0 - 1) ? lcp1 : -1));
    const int m1m = warp_max_u(max((ok0 && lane < B.depth1 - 1) ? lcp0 : -1, (ok1 && 32 + lane < B.depth1 - 1) ? lcp1 : -1));
    const int m2m = warp_max_u(max((ok0 && lane < B.depth2 - 1) ? lcp0 : -1, (ok1 && 32 + lane < B.depth2 - 1) ? lcp1 : -1));
    // find_match (matcher.rs:135-192)
    FmState st;
    st.max_len = kMatchMin - 1;
    st.exp_best = kMatchMin;
    st.best = -1;
    st.done = B.depth0 <= 0;
    fm_round(st, ok0 && lane < B.depth0, lcp0, kentry_exp(en), lane, lane,
             [&](int ml) { return ld_u32_unaligned(win, p0 + ml - 3) == ld_u32_unaligned(win, y + ml - 3); });
    if (n1 > 0)
        fm_round(st, ok1 && 32 + lane < B.depth0, lcp1, kentry_exp(en1), 32 + lane, lane,
                 [&](int ml) { return ld_u32_unaligned(win, p1 + ml - 3) == ld_u32_unaligned(win, y + ml - 3); });
    uint32_t mlen = 0, wexp = 0, word = 0;
    if (st.max_len >= kMatchMin && y + (uint32_t)st.max_len < B.sbuf_len && st.best >= 0) { // matcher.rs:183
        mlen = (uint32_t)st.max_len;
        const uint32_t bo0 = __shfl_sync(kFull, kentry_ord(en), st.best & 31), bo1 = __shfl_sync(kFull, kentry_ord(en1), st.best & 31);
        const uint32_t be0 = __shfl_sync(kFull, kentry_exp(en), st.best & 31), be1 = __shfl_sync(kFull, kentry_exp(en1), st.best & 31);
        word = st.best < 32 ? bo0 : bo1;
        wexp = st.best < 32 ? be0 : be1;
    }
    const uint32_t ord0 = bcast0(kentry_ord(en));
    const int last = ncand - 1;
    const uint32_t lo0 = __shfl_sync(kFull, kentry_ord(en), last >= 0 ? (last & 31) : 0), lo1 = __shfl_sync(kFull, kentry_ord(en1), last >= 0 ? (last & 31) : 0);
    const uint32_t ordmin = last < 0 ? kNone : (last < 32 ? lo0 : lo1); // kNone: no candidate
    // words updates of the two possible outcomes (lz.rs:208,233): words[hash2(s'-3)] = bytes s'-2, s'-1
    const uint32_t hw_lit = (b_m2 & 0x7fu) | ((b_m3 & 0x7fu) | is_alnum(b_m4) << 7) << 7; // s' = y+1
    const uint32_t wv_lit = b_m1 | (cur4 & 0xffu) << 8;
    uint32_t hw_m = 0, wv_m = 0;
    if (mlen > 0) { // s' = y + len
        const uint32_t t4 = ld_u32_unaligned(win, y + mlen - 5); // bytes s'-5 .. s'-2
        const uint32_t t2 = ld_u32_unaligned(win, y + mlen - 2); // bytes s'-2 .. s'+1
        const uint32_t x3 = (t4 >> 16) & 0xffu, x4 = (t4 >> 8) & 0xffu, x5 = t4 & 0xffu;
        hw_m = (x3 & 0x7fu) | ((x4 & 0x7fu) | is_alnum(x5) << 7) << 7;
        wv_m = t2 & 0xffffu;
    }
    ORZ_COST(200);
    if (B.dbg_pos && y - B.dbg_pos < 3u) { // (instrumented builds) the first four candidates: position, lcp | expected << 8
        uint32_t tv[9];
        tv[0] = y;
        for (int i = 0; i < 4; i++) {
            tv[1 + 2 * i] = __shfl_sync(kFull, ok0 ? p0 : 0u, i);
            tv[2 + 2 * i] = __shfl_sync(kFull, ok0 ? ((uint32_t)lcp0 | kentry_exp(en) << 8) : 0xffffffffu, i);
        }
        if (lane == 0) trace7(B.S, 5u, tv, 9);
    }
    // ---- publish: ext (all lanes), summary words 1.. (lane 0), fence, stamp
    uint32_t stampv = y;
    if (mode) {
        const uint32_t cur = bcast0(ld_vol_u32(&sum[0]));
        const uint32_t gen = (cur & 0x1ffffffu) == y ? cur >> 25 : 0u;
        stampv = y | ((gen + (mode == 1 ? 1u : 0u)) & 0x7fu) << 25;
        if (mode == 1) {
            if (lane == 0) *reinterpret_cast<volatile uint32_t *>(&sum[0]) = kNone;
            __threadfence(); // the stamp is gone before any word of the record changes
            __syncwarp();
        }
    }
#if ORZ_V7_TEST_TEAR
    if (mode == 0) { // the stamp first, then a pause in which the decide team and the resolver get to see the slot's old words under it
        if (lane == 0) *reinterpret_cast<volatile uint32_t *>(&sum[0]) = stampv;
        for (int i = 0; i < 6; i++) {
            __syncwarp();
            ORZ_POLL();
        }
    }
#endif
    const uint32_t xw0 = ok0 ? ((uint32_t)lcp0 | kentry_exp(en) << 8 | ((ord0 - kentry_ord(en)) & 0xfffu) << 16) : 0u;
    const uint32_t xw1 = (ok1 && 32 + lane < kExtCand) ? ((uint32_t)lcp1 | kentry_exp(en1) << 8 | ((ord0 - kentry_ord(en1)) & 0xfffu) << 16) : 0u;
    if (ok0) ext[2 + lane] = xw0;
    if (ok1 && 32 + lane < kExtCand) ext[2 + 32 + lane] = xw1;
    const uint32_t xcheck = extcheck7(y, F, ord0, xw0, xw1, lane);
    uint32_t *gx = B.gext + (size_t)slot * kExtCand;
    if (ok0) gx[lane] = p0;
    if (ok1 && 32 + lane < kExtCand) gx[32 + lane] = p1;
    if (lane == 0) {
        const uint32_t c1 = m1 < 0 ? 0u : (m1 > 255 ? 255u : (uint32_t)m1);
        const uint32_t c2 = m2 < 0 ? 0u : (m2 > 255 ? 255u : (uint32_t)m2);
        ext[0] = F;
        ext[1] = ord0;
        sum[1] = (c << 13 | e0) | (uint32_t)ncand << 21 | (incomplete ? 1u : 0u) << 27;
        sum[2] = h2 | (cur4 & 0xffffu) << 15;
        sum[3] = mlen | c1 << 8 | c2 << 16 | wexp << 24;
        sum[4] = word;
        sum[5] = ordmin;
        sum[6] = F;
        sum[7] = hw_lit | wv_lit << 15;
        sum[8] = hw_m | wv_m << 15;
        sum[9] = (uint32_t)(mf < 0 ? 0 : mf) | (uint32_t)(m1m < 0 ? 0 : m1m) << 8 | (uint32_t)(m2m < 0 ? 0 : m2m) << 16;
        sum[10] = sumcheck7(y, (c << 13 | e0) | (uint32_t)ncand << 21 | (incomplete ? 1u : 0u) << 27, h2 | (cur4 & 0xffffu) << 15, mlen | c1 << 8 | c2 << 16 | wexp << 24, word, ordmin, F,
                            hw_lit | wv_lit << 15, hw_m | wv_m << 15, (uint32_t)(mf < 0 ? 0 : mf) | (uint32_t)(m1m < 0 ? 0 : m1m) << 8 | (uint32_t)(m2m < 0 ? 0 : m2m) << 16);
        sum[11] = xcheck;
#if ORZ_V7_PROV
        if (B.S->prov) { // provenance of the build, by absolute input position (three uint4 per position: build, build, last repair)
            const size_t pi = (size_t)B.prov_base + (y - kPrematch);
            if (y >= kPrematch && pi < (size_t)B.S->prov_cap) {
                B.S->prov[4 * pi] = make_uint4(F, cnt, r | (uint32_t)mode << 8 | (uint32_t)n0 << 16 | (uint32_t)n1 << 24, gc);
                B.S->prov[4 * pi + 1] = make_uint4((uint32_t)ncand | (incomplete ? 1u : 0u) << 8 | ((uint32_t)st.best & 0xffu) << 16, mlen | c1 << 8 | c2 << 16 | wexp << 24, en.x, en.y);
            }
        }
#endif
        if (B.dbg_pos && y - B.dbg_pos < 3u) {
            const uint32_t tv[14] = {y, F, cnt, r, (uint32_t)n0 | (uint32_t)n1 << 8 | (uint32_t)ncand << 16 | (incomplete ? 1u : 0u) << 24, gc, mlen | c1 << 8 | c2 << 16 | wexp << 24,
                                     (uint32_t)mode, word, (uint32_t)st.best, st.exp_best, ord0, ordmin, avail};
            trace7(B.S, 1u, tv, 14);
        }
    }
    // EVERY lane fences its own stores (ext words through DSMEM, candidate positions in HBM) before the stamp goes out.  A warp
    // barrier followed by lane 0's release store is not enough on the hardware: measured on a B200, about one 100 MB run in eight
    // had a repair read ext words of the slot's previous record right after a just-in-time build (the stamp had overtaken them).
    fence_cluster7();
    __syncwarp();
    if (lane == 0) st_release_u32(&sum[0], stampv);
}

// Window bytes for repairs: from the shared-memory ring where it is guaranteed stable (the feeder only moves when the parse
// position does, and a repair holds it still), from HBM otherwise.
struct ByteSrc7 {
    const uint8_t *win;
    const uint32_t *ring;
    uint32_t lo, hi; // ring-safe byte range [lo, hi)
    __device__ __forceinline__ uint32_t u32(uint32_t p) const {
        if (p >= lo && p + 8 <= hi) {
            const uint32_t w0 = ring[(p >> 2) & (kInRing / 4 - 1)], w1 = ring[((p >> 2) + 1) & (kInRing / 4 - 1)];
            return __funnelshift_r(w0, w1, (p & 3u) * 8u);
        }
        return ld_u32_unaligned(win, p);
    }
    __device__ __forceinline__ uint32_t ring_u32(uint32_t p) const {
        const uint32_t w0 = ring[(p >> 2) & (kInRing / 4 - 1)], w1 = ring[((p >> 2) + 1) & (kInRing / 4 - 1)];
        return __funnelshift_r(w0, w1, (p & 3u) * 8u);
    }
    // the same for the whole warp at once (uniform a, b): lane l compares bytes 8l .. 8l+7
    __device__ __forceinline__ int lcp_warp(uint32_t a, uint32_t b, int lane) const {
        if (a >= lo && b >= lo && a + 264 <= hi && b + 264 <= hi) {
            const uint32_t o = 8u * (uint32_t)lane;
            const uint32_t d0 = ring_u32(a + o) ^ ring_u32(b + o), d1 = ring_u32(a + o + 4) ^ ring_u32(b + o + 4);
            const int mine = d0 ? ((__ffs(d0) - 1) >> 3) : (d1 ? 4 + ((__ffs(d1) - 1) >> 3) : 8);
            const unsigned mm = __ballot_sync(kFull, lane < kMatchMax / 8 && mine < 8);
            if (!mm) return kMatchMax;
            const int l = __ffs(mm) - 1;
            return 8 * l + __shfl_sync(kFull, mine, l);
        }
        return lcp(a, b);
    }
    // mem.rs:41-51 restated byte-exact: longest common prefix of [a..) and [b..), capped at 240
    __device__ __forceinline__ int lcp(uint32_t a, uint32_t b) const {
        if (a >= lo && b >= lo && a + 248 <= hi && b + 248 <= hi) { // both ranges sit in the ring
            for (int i = 0; i < kMatchMax; i += 4) {
                const uint32_t x = ring_u32(a + i) ^ ring_u32(b + i);
                if (x) return i + ((__ffs(x) - 1) >> 3);
            }
            return kMatchMax;
        }
        return lcp_win_win(win, a, b);
    }
};

// REPAIR without draining.  The record of y was built against the state below its frontier F_ext; since then steps with y's key
// may have been decided (they sit in the step log, committed or not), and candidates may have left the ring.  The exact candidate
// sequence at decision time T is  [steps in [F_ext, T) with y's key, newest first]  followed by  [the record's own candidates that
// are still alive]; matcher.rs:135-228 is re-resolved over that sequence and the summary rewritten with frontier T.
// Returns 0 if the record has to be rebuilt from HBM instead (incomplete record, log does not reach back to F_ext, > 32 recent steps),
// else how it was repaired: 1 = nothing to do (the key-hash table aliased), 2 = one newer node took the first place, 3 = re-resolved.
__device__ __forceinline__ int repair_record7(ParseShared7 &shl, const Build7 &B, const ByteSrc7 &src, uint32_t y, uint32_t slot, uint32_t T,
                                               uint32_t qhead, int lane, uint32_t *dbg4 = nullptr) {
    uint32_t *rs = shl.sum[slot];
    const uint32_t *ex = shl.ext[slot];
    // the record must be whole before anything is derived from it (see sumcheck7): wait for late words, give up after a while
    uint32_t w1 = 0;
    for (int tries = 0;; tries++) {
        w1 = ld_vol_u32(rs + 1);
        const uint32_t n = (w1 >> 21) & 0x3fu;
        const uint32_t c0 = (uint32_t)lane < n ? ld_vol_u32(ex + 2 + lane) : 0u;
        const uint32_t c1 = (32u + (uint32_t)lane < n && 32 + lane < kExtCand) ? ld_vol_u32(ex + 34 + lane) : 0u;
        // (the ext checksum holds a warp collective: every lane computes it, whatever its own view of the summary words says)
        const uint32_t xc = extcheck7(y, ld_vol_u32(ex), ld_vol_u32(ex + 1), c0, c1, lane);
        const uint32_t sc = sumcheck7(y, w1, ld_vol_u32(rs + 2), ld_vol_u32(rs + 3), ld_vol_u32(rs + 4), ld_vol_u32(rs + 5), ld_vol_u32(rs + 6), ld_vol_u32(rs + 7), ld_vol_u32(rs + 8),
                                      ld_vol_u32(rs + 9));
        const bool ok = ORZ_V7_MUT_NOCHECK || (sc == ld_vol_u32(rs + 10) && xc == ld_vol_u32(rs + 11));
        if (__all_sync(kFull, ok)) break;
        if (tries == 0 && lane == 0) atomicAdd(&shl.torn, 0x10000u);
        if (tries > 200000) return 0; // (never seen: the record is then rebuilt from HBM)
        ORZ_POLL();
    }
    if ((w1 >> 27) & 1u) return 0;
    const uint32_t key = w1 & 0x1fffffu, nold = (w1 >> 21) & 0x3fu, c = key >> 13;
    const uint32_t F_ext = ex[0], ord0 = ex[1];
    const uint32_t bc = shl.bcnt[c];
    const unsigned lt_mask = (1u << lane) - 1u;
    // ---- 1. steps with this key decided since the record's frontier, newest first
    uint32_t k = 0, newer_c = 0; // newer_c: steps of bucket c seen so far (all newer than the entry being looked at)
    bool reached = false;
    for (int r = 0; r < 14 && !reached; r++) {
        const int idx = (int)qhead - 1 - (r * 32 + lane);
        uint4 e = make_uint4(0, 0, 0, 0);
        if (idx >= 0) e = shl.queue[idx % kQueue7];
        const uint32_t epos = e.x & 0x1ffffffu;
        const bool inrange = idx >= 0 && epos >= F_ext;
        const bool hit = inrange && (e.z & 0x1fffffu) == key;
        const unsigned hm = __ballot_sync(kFull, hit);
        const unsigned cm = __ballot_sync(kFull, idx >= 0 && ((e.z >> 13) & 0xffu) == c);
        const uint32_t slot = k + (uint32_t)__popc(hm & lt_mask);
        if (hit && slot < 32u) {
            shl.rc_pos[slot] = epos;
            shl.rc_exp[slot] = e.z >> 21;
            shl.rc_ord[slot] = (bc - 1 - (newer_c + (uint32_t)__popc(cm & lt_mask))) & kOrdMask; // its insert's ordinal in the bucket
        }
        k += (uint32_t)__popc(hm);
        newer_c += (uint32_t)__popc(cm);
        if (__ballot_sync(kFull, !inrange)) reached = true;
    }
    if (dbg4) {
        dbg4[0] = w1;
        dbg4[1] = F_ext;
        dbg4[2] = k | (reached ? 1u : 0u) << 8 | bc << 12;
        dbg4[3] = 0xffffffffu;
    }
    if (!reached || k > 32u) return 0;
    if (k == 0) { // the key-hash table aliased: nothing decided since F_ext has this key
        const uint32_t om = rs[5];
        if (om == kNone || ((bc - 1 - om) & kOrdMask) <= (uint32_t)(kRing - 1)) {
            __syncwarp();
            if (lane == 0) rs[6] = T;
            __syncwarp();
            return 1;
        }
    }
    __syncwarp();
    if (k == 1) {
        // ---- exactly one newer node q1 of this key, every candidate of the record still alive: q1 is tried first (matcher.rs:149).
        // If it ends the search at once (lcp 240, or longer than its own expected length: matcher.rs:162-167) or no other
        // candidate is longer, the result is q1 — no need to go through the sequence.
        const uint32_t om = rs[5];
        const bool alive = om == kNone || ((bc - 1 - om) & kOrdMask) <= (uint32_t)(kRing - 1);
        const uint32_t q1 = shl.rc_pos[0], e1 = shl.rc_exp[0], o1 = shl.rc_ord[0], mx = rs[9];
        const uint32_t l1 = (uint32_t)src.lcp_warp(q1, y, lane);
        const uint32_t fmmax = B.depth0 >= 2 ? (mx & 0xffu) : 0u;
        if (alive && B.depth0 >= 1 && l1 >= (uint32_t)kMatchMin && (l1 == (uint32_t)kMatchMax || (e1 > 0 && l1 > e1) || l1 >= fmmax)) {
            const uint32_t mlen = y + l1 < B.sbuf_len ? l1 : 0u; // matcher.rs:183
            const uint32_t z1 = B.depth1 >= 1 ? (B.depth1 >= 2 && ((mx >> 8) & 0xffu) > l1 ? (mx >> 8) & 0xffu : l1) : 0u;
            const uint32_t z2 = B.depth2 >= 1 ? (B.depth2 >= 2 && ((mx >> 16) & 0xffu) > l1 ? (mx >> 16) & 0xffu : l1) : 0u;
            uint32_t hw_m = 0, wv_m = 0;
            if (mlen > 0) { // lz.rs:208
                const uint32_t t4 = src.u32(y + mlen - 5), t2 = src.u32(y + mlen - 2);
                const uint32_t x3 = (t4 >> 16) & 0xffu, x4 = (t4 >> 8) & 0xffu, x5 = t4 & 0xffu;
                hw_m = (x3 & 0x7fu) | ((x4 & 0x7fu) | is_alnum(x5) << 7) << 7;
                wv_m = t2 & 0xffffu;
            }
            __syncwarp();
            if (lane == 0) {
                rs[3] = mlen | (z1 > 255u ? 255u : z1) << 8 | (z2 > 255u ? 255u : z2) << 16 | (mlen ? e1 : 0u) << 24;
                rs[4] = o1;
                if (om == kNone) rs[5] = o1;
                rs[6] = T;
                rs[8] = hw_m | wv_m << 15;
            }
            __syncwarp();
            return 2;
        }
    }
    // ---- 2. merged candidate sequence, 32 per round: lanes < k hold the recent steps (round 0), the others the record's candidates
    int dmax = B.depth0 > B.depth1 ? B.depth0 : B.depth1;
    dmax = dmax > B.depth2 ? dmax : B.depth2;
    FmState st;
    st.max_len = kMatchMin - 1;
    st.exp_best = kMatchMin;
    st.best = -1;
    st.done = B.depth0 <= 0;
    int m1 = -1, m2 = -1;
    uint32_t best_ord = 0, best_exp = 0, last_ord = kNone;
    uint32_t *gx = B.gext + (size_t)slot * kExtCand;
    for (int round = 0; round < 2; round++) {
        const int m = round * 32 + lane;
        bool valid = false;
        int lcp = -1;
        uint32_t exp = 0, ord = 0, cpos = 0;
        bool recent = false;
        if (m < (int)k) {
            recent = true;
            cpos = shl.rc_pos[m];
            exp = shl.rc_exp[m];
            ord = shl.rc_ord[m];
            valid = m < dmax;
            if (valid) lcp = src.lcp(cpos, y);
        } else {
            const int j = m - (int)k;
            if (j < (int)nold && m < dmax) {
                const uint32_t w = ex[2 + j];
                lcp = (int)(w & 0xffu);
                exp = (w >> 8) & 0xffu;
                ord = (ord0 - ((w >> 16) & 0xfffu)) & kOrdMask;
                valid = ((bc - 1 - ord) & kOrdMask) <= (uint32_t)(kRing - 1); // still among the bucket's last 4094 inserts
            }
        }
        const unsigned vm = __ballot_sync(kFull, valid);
        const int nv = vm == 0xffffffffu ? 32 : (__ffs(~vm) - 1); // alive candidates form a prefix
        if (dbg4 && round == 0) dbg4[3] = (uint32_t)nv | (uint32_t)dmax << 8 | nold << 16 | (uint32_t)(__shfl_sync(kFull, lcp, 0) & 0xff) << 24;
        valid = lane < nv;
        m1 = max(m1, warp_max_u((valid && m < B.depth1) ? lcp : -1));
        m2 = max(m2, warp_max_u((valid && m < B.depth2) ? lcp : -1));
        const int before = st.best;
        fm_round(st, valid && m < B.depth0, lcp, exp, m, lane, [&](int ml) {
            const uint32_t p = recent ? cpos : ldcg_u32(&gx[m - (int)k]);
            return src.u32(p + ml - 3) == src.u32(y + ml - 3);
        });
        if (st.best != before && st.best >= round * 32) {
            best_ord = __shfl_sync(kFull, ord, st.best & 31);
            best_exp = __shfl_sync(kFull, exp, st.best & 31);
        }
        if (nv > 0) last_ord = __shfl_sync(kFull, ord, nv - 1);
        if (nv < 32) break;
    }
    uint32_t mlen = 0, wexp = 0, word = 0;
    if (st.max_len >= kMatchMin && y + (uint32_t)st.max_len < B.sbuf_len && st.best >= 0) { // matcher.rs:183
        mlen = (uint32_t)st.max_len;
        word = best_ord;
        wexp = best_exp;
    }
    uint32_t hw_m = 0, wv_m = 0;
    if (mlen > 0) { // lz.rs:208: words[hash2(s'-3)] = bytes s'-2, s'-1 with s' = y + len
        const uint32_t t4 = src.u32(y + mlen - 5), t2 = src.u32(y + mlen - 2);
        const uint32_t x3 = (t4 >> 16) & 0xffu, x4 = (t4 >> 8) & 0xffu, x5 = t4 & 0xffu;
        hw_m = (x3 & 0x7fu) | ((x4 & 0x7fu) | is_alnum(x5) << 7) << 7;
        wv_m = t2 & 0xffffu;
    }
    __syncwarp();
    if (lane == 0) {
        const uint32_t c1 = m1 < 0 ? 0u : (m1 > 255 ? 255u : (uint32_t)m1);
        const uint32_t c2 = m2 < 0 ? 0u : (m2 > 255 ? 255u : (uint32_t)m2);
        rs[3] = mlen | c1 << 8 | c2 << 16 | wexp << 24;
        rs[4] = word;
        rs[5] = last_ord;
        rs[6] = T;
        rs[8] = hw_m | wv_m << 15;
    }
    __syncwarp();
    return 3;
}

// table index of a key (21 bits) or a words index (15 bits) for the verification of a super-window
__device__ __forceinline__ uint32_t vhash7(uint32_t v) { return ((v ^ (v >> 9)) * 0x9E3779B1u) >> 24; }

// Record slots are private to builders: builder b = y % NBW owns SL = 512 / NBW slots and is the only writer of the record of y,
// (y % NBW) * SL + (y / NBW) % SL.  A builder handles its positions in increasing order, so a build that finishes late (the parse
// already went past y) can only overwrite that builder's own, older record — never a newer record of another builder.
template <int NBW>
__device__ __forceinline__ uint32_t slot7(uint32_t y) {
    constexpr uint32_t SL = (uint32_t)kRing7 / (uint32_t)NBW;
    return (y % (uint32_t)NBW) * SL + (y / (uint32_t)NBW) % SL;
}

// One block of the decide team: lane l takes the decision of position base + l against the tables as they stand (lz.rs:131-236),
// as if it were the next step, and leaves the chase word and the step's data for the resolver.  `block0`: lane 0 of this block is
// the first step of the super-window (records that are exact only for a first step are usable there).
template <int NBW>
__device__ __forceinline__ void decide_block7(ParseShared7 &shl, uint32_t round_base, uint32_t base, bool block0, uint32_t sbuf_len, int lane, uint2 *pk_out,
                                              uint32_t (*ent_out)[8], EncStream *S, uint32_t dbg_pos) {
    const uint32_t x = base + (uint32_t)lane;
    const uint32_t *rs = shl.sum[slot7<NBW>(x)];
    // The stamp is the builder's last store (release); the first 16 bytes are read in one shared-memory transaction, the rest
    // afterwards, so a matching stamp covers every field.
    const uint4 a = lds_volatile_v4(rs);
    const uint4 b = lds_volatile_v4(rs + 4);
    const uint4 c4 = lds_volatile_v4(rs + 8); // words update of a match, maxima without the last candidate, the two checksums
    const uint32_t wmat = c4.x;
    const uint32_t stamp2 = ld_vol_u32(rs); // (a rebuild in place takes the stamp away first and returns it with another generation)
    const bool sum_ok = ORZ_V7_MUT_NOCHECK || sumcheck7(x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c4.x, c4.y) == c4.z;
    const bool ready = (a.x & 0x1ffffffu) == x && a.x == stamp2 && sum_ok;
    if ((a.x & 0x1ffffffu) == x && a.x == stamp2 && !sum_ok) atomicAdd(&shl.torn, 1u); // (counted: the stamp was there, the words were not)
    const uint32_t key = a.y & 0x1fffffu, c = key >> 13;
    const bool incomplete = (a.y >> 27) & 1u;
    const uint32_t h2 = a.z & 0x7fffu, b2 = a.z >> 15;
    const uint32_t mlen = a.w & 0xffu, wexp = a.w >> 24;
    const uint32_t word = b.x, ordmin = b.y, F = b.z, wlit = b.w;
    const uint32_t lwe = shl.words[h2];                            // lz.rs:132
    const uint32_t bc = ld_vol_u32(&shl.bcnt[c]);
    const uint32_t kh = (key ^ (key >> 9)) & 4095u;
    const uint32_t lk = ld_vol_u32(&shl.lastkey[kh]);
    // exactness of the record at decision time (see the header): no step with this key since F, oldest candidate alive.
    // `room`: this many further inserts into the bucket keep the record's oldest candidate alive (checked at verification).
    // flag bits: 1 = stale (repair), 2 = not built yet.  (Branch-free on purpose: a warp that is alone on its SM
    // sub-partition pays a pipeline refill for every divergent branch.)
    const int room = ordmin == kNone ? 64 : (int)(kRing - 1) - (int)((bc - 1 - ordmin) & kOrdMask);
    const uint32_t fown = !ready ? 2u : ((lk >= F || incomplete || room < 0) ? 1u : 0u);
    // a stale record far enough ahead is handed back to its builder: rebuilt against the present frontier it is exact again by the
    // time the parse gets there (if the step that made it stale is below the frontier by now; the parse repairs the near ones itself)
    if (ready && (fown & 1u) && x >= round_base + kRefreshAhead && x <= sbuf_len + 1u && (incomplete || room < 0 || lk < ld_vol_u32(&shl.fpos))) {
        const uint32_t slot = slot7<NBW>(x), w = slot >> 5, bit = 1u << (slot & 31u);
        if (!((ld_vol_u32(&shl.req[w]) | ld_vol_u32(&shl.busy[w])) & bit)) atomicOr(&shl.req[w], bit);
    }
    // neighbours: lane+1 supplies the depth1 lazy maximum, lane+2 the depth2 one (lz.rs:154-169)
    const uint32_t room5 = room < 0 ? 0u : (room > 31 ? 31u : (uint32_t)room);
    const uint32_t nb = ((a.w >> 8) & 0xffffu) | fown << 16 | room5 << 19;
    const uint32_t nb1 = __shfl_down_sync(kFull, nb, 1), nb2 = __shfl_down_sync(kFull, nb, 2);
    const uint32_t key1 = __shfl_down_sync(kFull, key, 1), key2 = __shfl_down_sync(kFull, key, 2);
    const uint32_t lzn1 = nb1 & 0xffu, lzn2 = (nb2 >> 8) & 0xffu;
    const bool lzneed = mlen > 0 && mlen < (uint32_t)(kMatchMax / 2);   // lz.rs:151
    const uint32_t lwm = lwe == b2 ? 1u : 0u;                       // lz.rs:133
    const uint32_t ro0 = (bc - 1 - word) & kOrdMask;                // reduced offset with no further insert before it (matcher.rs:185)
    const uint32_t rb = ro0 < 510u ? 1u : 0u;                       // robitlen < 8  <=>  reduced offset < 510
    // lz.rs:146-236, spelled out without branches
    const uint32_t l1 = mlen + 1 + rb, l2 = l1 - lwm;               // lz.rs:152-153
    const bool lazy1 = lzneed && lzn1 >= l1;
    const bool lazy2 = lzneed && !lazy1 && lzn2 >= l2;
    const bool is_match = mlen > 0 && !lazy1 && !lazy2;
    const bool is_word = !is_match && x + 1 < sbuf_len && !lazy1 && lwm; // lz.rs:218
    // An insert of this super-window into the bucket could flip robitlen < 8 (the reduced offset only grows: 1 -> 0).  The
    // flag only enters the two lazy thresholds (lz.rs:152-153; the emitted code comes from the committer's own ordinals), so
    // the step is sensitive only if the decision taken with the other value differs.
#if ORZ_V7_EDGE_EXACT
    const bool lazy1f = lzneed && lzn1 >= mlen + 1;
    const bool lazy2f = lzneed && !lazy1f && lzn2 >= mlen + 1 - lwm;
    const bool rbsens = lzneed && ro0 < 510u && ro0 + 31u >= 510u && (lazy1f != lazy1 || lazy2f != lazy2);
#else
    const bool rbsens = mlen > 0 && ro0 < 510u && ro0 + 31u >= 510u;
#endif
    const uint32_t ty = is_match ? kTyMatch : (is_word ? kTyWord : kTyLit);
    const uint32_t adv = is_match ? mlen : (is_word ? 2u : 1u);
    // a lane either takes a step or stops the chase: code 1 = a record must be repaired first, 2 = a record is not built yet,
    // 3 = next block, 4 = end of the super-window.  The records of x+1 / x+2 only matter if the step evaluates lazy matches; a
    // robitlen < 8 that an insert could flip is exact only for the FIRST step of a super-window.
    const uint32_t m = (fown & 3u) | (lzneed ? (((nb1 | nb2) >> 16) & 3u) : 0u);
    // a tight record of a lazy neighbour: exact for a step that has at most `room` steps of the super-window before it (whatever
    // bucket they insert into) — also checked at verification
    const uint32_t r1 = (nb1 >> 19) & 31u, r2 = (nb2 >> 19) & 31u;
    const uint32_t nbroom = lzneed ? (r1 < r2 ? r1 : r2) : 31u;
    const bool first = block0 && lane == 0;
    const uint32_t scode = (m & 2u) ? 2u : (m & 1u) ? 1u : ((!first && rbsens) || x >= sbuf_len) ? 4u : lane >= 30 ? 3u : 0u;
    // (a repair stop also says which of the three records are stale: bit 0 own, 1 x+1, 2 x+2)
    const uint32_t smask = (fown & 1u) | ((nb1 >> 16) & 1u) << 1 | ((nb2 >> 16) & 1u) << 2;
    // The chase of the step chain through this block, from every lane at once (pointer doubling, 5 rounds cover 32 hops): a step lane
    // points at lane + advance, a stop lane at itself.  Afterwards lane l knows the lanes a chase entering at l steps on and where it
    // ends: a position past the block (>= 32, the true advance of the last step) or the stop lane it runs into.
    uint32_t nxt = scode ? (uint32_t)lane : (uint32_t)lane + adv;
    uint32_t msk = scode ? 0u : 1u << lane;
#pragma unroll
    for (int r = 0; r < 5; r++) {
        const uint32_t m2 = __shfl_sync(kFull, msk, (int)(nxt & 31u)), n2 = __shfl_sync(kFull, nxt, (int)(nxt & 31u));
        if (nxt < 32u) {
            msk |= m2;
            nxt = n2;
        }
    }
    const uint32_t sinfo = __shfl_sync(kFull, scode | smask << 4, (int)(nxt & 31u));
    // x = lanes stepped on; y = end (lane or position relative to the block) | stop code << 12 | stale mask << 16
    const uint2 pk = make_uint2(msk, nxt | (nxt < 32u ? sinfo << 12 : 0u));
    const uint32_t wu = ty == kTyWord ? 0xffffffffu : (ty == kTyMatch ? wmat : wlit); // lz.rs:208,233
    if (dbg_pos && x - dbg_pos < 3u) {
        const uint32_t tv[10] = {x, round_base, fown | scode << 4 | ty << 8 | (lzneed ? 1u : 0u) << 12, lk, F, mlen | lzn1 << 8 | lzn2 << 16, lwm | rb << 1, ro0, a.x, bc};
        trace7(S, 4u, tv, 10);
    }
    pk_out[lane] = pk;
    *reinterpret_cast<uint4 *>(ent_out[lane]) = make_uint4(x, ty | (lzneed ? 4u : 0u) | room5 << 3 | adv << 8 | mlen << 16 | wexp << 24, key1 | (b2 & 0xffu) << 21, key2 | nbroom << 21);
    *reinterpret_cast<uint4 *>(ent_out[lane] + 4) = make_uint4(key, wu, h2 | lwe << 15, word);
}

// NW warps per CTA, K CTAs per stream (a thread-block cluster).  CTA 0: warp 0 resolver, warp 1 committer, the warps that
// do not share the resolver's SM sub-partition build; CTAs 1..K-1: all warps build.
template <int NW, int K, bool STATS>
__global__ void __launch_bounds__(NW * 32) orz_parse7_kernel(EncStream *streams, int buf) {
    constexpr int NT = NW * 32;
    static_assert(NW % 4 == 0 && NW >= 8, "stage A v7: warps per CTA");
    // Builders in CTA 0: none when the stream owns a cluster — measured on a B200, every busy warp on the resolver's SM slows the
    // resolver (shared LSU / shared-memory pipeline): 1225 cycles per step with 9 builders beside it.  A single-CTA launch (K = 1)
    // has to keep them: warps >= 4 that do not share the resolver's SM sub-partition.
    constexpr int NB0 = K > 1 ? 0 : NW - 4 - kTeam7 - ((NW - 4 - kTeam7) / 4); // (NW = 16: warps 9, 10, 11, 13, 14, 15)
    constexpr int NBW = NB0 + (K - 1) * NW;    // builder warps of the cluster
    constexpr uint32_t kLeadMax = ((uint32_t)kRing7 / NBW) * NBW - 40u;
    constexpr uint32_t kLead = kLeadMax < kLeadTarget7 ? kLeadMax : kLeadTarget7; // builders stay this close to the parse position
    EncStream *S = &streams[blockIdx.x / K];
    const unsigned rank = K > 1 ? cluster_cta_rank() : 0u;
    ORZ_DYN_SHARED(smem_raw);
    ParseShared7 &shl = *reinterpret_cast<ParseShared7 *>(smem_raw);
    ParseShared7 &sh = K > 1 ? *map_to_cta0(&shl) : shl; // CTA 0's shared memory (DSMEM for the other CTAs)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool active = S->status == 0 && S->parse_done == 0; // uniform over the cluster (the status word is not written before the final barrier)

    if (active && rank == 0) {
        for (int i = tid; i < kNumCtx; i += NT) shl.cnt[i] = 0;
        for (int i = tid; i < kNumBuckets; i += NT) shl.bcnt[i] = shl.lcnt[i] = shl.icnt[i] = S->bcnt[i];
        for (int i = tid; i < 4096; i += NT) shl.lastkey[i] = 0;
        for (int i = tid; i < 512; i += NT) shl.vtab[i] = 0;
        for (int i = tid; i < kRing7; i += NT) shl.sum[i][0] = kNone;
        for (int i = tid; i < kRing7 / 32; i += NT) shl.req[i] = shl.busy[i] = 0;
        for (int i = tid; i < 32768 / 8; i += NT) reinterpret_cast<uint4 *>(shl.words)[i] = reinterpret_cast<const uint4 *>(S->words)[i];
        if (tid == 0) {
            shl.spos = S->spos;
            shl.lead_pos = S->spos;
            shl.fpos = S->spos;
            shl.qhead = 0;
            shl.qdone = 0;
            shl.qitem = 0;
            shl.done = 0;
            shl.team_base = S->spos;
            shl.torn = 0;
            // the feeder starts 2048 bytes behind the parse position (16-byte aligned, inside the allocated window)
            uint32_t f0 = (S->spos - 2048u) & ~15u;
            const uint32_t wlo = (S->win_lo + 15u) & ~15u;
            if (S->win_lo > 0 && f0 < wlo) f0 = wlo;
            shl.fed_lo0 = f0;
            shl.fed_hi = f0;
        }
    }
    __syncthreads();
    if (K > 1) cluster_barrier(); // CTA 0's shared memory is initialised before any remote builder looks at it
    if (active) {
        const uint32_t sbuf_len = S->sbuf_len;
        Build7 B;
        B.S = S;
        B.prov_base = (uint32_t)(S->in_pos - (S->sbuf_len - kPrematch));
        B.dbg_pos = ((STATS || ORZ_V7_TRACE_ALWAYS) && S->dbg_pos != 0u && S->in_pos == S->dbg_inpos) ? S->dbg_pos : 0u;
        B.win = S->win;
        B.klist = S->klist;
        B.kcnt = S->kcnt;
        B.gext = S->gext;
        B.sbuf_len = sbuf_len;
        B.depth0 = (int)S->depth0;
        B.depth1 = (int)S->depth1;
        B.depth2 = (int)S->depth2;
        volatile uint32_t *v_spos = &sh.spos;
        volatile uint32_t *v_done = &sh.done;

        if (rank == 0 && warp == 0) {
            // ============================================================== resolver
            volatile uint32_t *v_qhead = &shl.qhead;
            volatile uint32_t *v_qdone = &shl.qdone;
            volatile uint32_t *v_qitem = &shl.qitem;
            const uint32_t items_cap = S->items_cap < kChunkItems ? S->items_cap : kChunkItems;
            uint32_t spos = S->spos;
            uint32_t after_literal = S->after_literal;
            uint32_t n_items = 0, qhead = 0;
            unsigned long long n_windows = 0, n_wait = 0, n_repair = 0, n_rep_rec = 0, n_rep_hard = 0, n_cut = 0, n_handed = 0;
            long long cyc_wait = 0, cyc_walk = 0, cyc_commit = 0, cyc_repair = 0, cyc_load = 0;
            unsigned long long n_stopc[5] = {0, 0, 0, 0, 0}, n_how[4] = {0, 0, 0, 0}, n_rwait = 0;
            unsigned long long lag_steps = 0, lag_pos = 0, probe_alu = 0, probe_shfl = 0, probe_lds = 0, probe_n = 0;
            const long long t_begin = STATS ? clock64() : 0;
            const unsigned lt_mask = (1u << lane) - 1u;
            int wd_code = 0;
            unsigned idle = 0;
            uint32_t qlimit = (uint32_t)kQueue7 - 8u; // queue entries below this index are free (consumers re-read when it is reached)
            const uint32_t last_pos = sbuf_len + 1; // the last position that has a record
            while (spos < sbuf_len && n_items < items_cap) {
                // ======== one super-window: up to 32 steps, gathered by passes over 30 positions each, then verified and handed over
                // with lane = step.  Every decision of the super-window is taken against the tables as they stand at its start;
                // the verification below finds the first step that an earlier step of the same super-window would have changed.
                const long long t0 = STATS ? clock64() : 0;
                const uint32_t max_steps = items_cap - n_items < 32u ? items_cap - n_items : 32u;
                uint32_t nst = 0, base = spos, rmask = 0;
                int stop = 0;
                // ---- the decide team takes the decision of the next kTeam7 * 30 positions, one block per warp, all at once
                if (lane == 0) *reinterpret_cast<volatile uint32_t *>(&shl.team_base) = spos;
                __threadfence_block(); // (bar.arrive itself orders nothing)
                __syncwarp();
                bar_arrive7(1, kTeamThreads); // go: the team warps are parked at barrier 1
                bar_sync7(2, kTeamThreads);   // done: every team warp arrives at barrier 2 when its block is written
                const long long tq = STATS ? clock64() : 0;
                if (STATS) cyc_load += tq - t0;
                if (STATS && (n_windows & 63u) == 0) {
                    // latency probes of this warp under the launch's real load: 32 dependent adds, 16 dependent shuffles, 16 dependent
                    // shared loads (pointer chase through sidx, filled with a fixed permutation of itself below)
                    uint32_t acc = (uint32_t)lane;
                    const long long p0 = clock64();
#pragma unroll
                    for (int i = 0; i < 32; i++) acc = acc * 3u + (uint32_t)i;
                    const long long p1 = clock64();
                    uint32_t sv = acc;
#pragma unroll
                    for (int i = 0; i < 16; i++) sv = __shfl_sync(kFull, sv, (int)(sv & 31u));
                    const long long p2 = clock64();
                    uint32_t li = (uint32_t)lane;
#pragma unroll
                    for (int i = 0; i < 16; i++) li = ld_vol_u32(&shl.lastkey[(li * 33u + 7u) & 4095u]) & 4095u;
                    const long long p3 = clock64();
                    probe_alu += (unsigned long long)(p1 - p0);
                    probe_shfl += (unsigned long long)(p2 - p1);
                    probe_lds += (unsigned long long)(p3 - p2);
                    probe_n++;
                    if (sv == 0xdeadbeefu && li == 0x12345u) shl.rc_pos[0] = acc; // (keeps the probes alive)
                }
                const long long tq2 = STATS ? clock64() : 0;
                // ---- chase the step chain through the blocks (lz.rs:131-236): one shuffle per step, nothing else on the chain
                uint32_t rel = 0; // position of the chase relative to spos
                for (uint32_t t = 0, cur = 0;;) {
                    const uint2 pv = lds_volatile_v2(&shl.pk_all[t][cur]); // (one broadcast load: the team chased the block from every lane)
                    uint32_t path = pv.x;
                    int code = (int)((pv.y >> 12) & 0xfu);
                    uint32_t np = (uint32_t)__popc(path);
                    if (nst + np > max_steps) { // keep only the steps that still fit (the chunk's item limit, lz.rs:131, or 32)
                        uint32_t keep = path;
                        for (uint32_t i = np; nst + i > max_steps; i--) keep &= ~(1u << (31 - __clz(keep)));
                        cur = (uint32_t)(__ffs(path & ~keep) - 1); // the first step that no longer fits
                        path = keep;
                        np = (uint32_t)__popc(path);
                        code = 4;
                    } else {
                        cur = pv.y & 0xfffu; // the stop lane, or the position past the block the last step lands on
                        rmask = (pv.y >> 16) & 7u;
                    }
                    if ((path >> lane) & 1u) shl.sidx[nst + (uint32_t)__popc(path & lt_mask)] = t * 32u + (uint32_t)lane;
                    nst += np;
                    rel = 30u * t + cur;
                    if (code == 1 || code == 2 || code == 4) {
                        stop = code;
                        break;
                    }
                    t = rel / 30u;
                    cur = rel - 30u * t;
                    if (t >= (uint32_t)kTeam7 || nst >= max_steps || spos + rel >= sbuf_len) break;
                }
                base = spos + rel;
                // let the builders move on: their lead is counted from here.  (Not `spos`: a verification cut may fall back behind this
                // position and take another path, so builders must not skip positions on account of it.)
                if (lane == 0) *reinterpret_cast<volatile uint32_t *>(&shl.lead_pos) = base;
                const long long t1 = STATS ? clock64() : 0;
                if (STATS) {
                    cyc_walk += t1 - tq2;
                    n_windows++;
                }
                if (nst == 0 && stop == 2) { // the next step's records are not there yet
                    if (STATS) {
                        n_wait++;
                        cyc_wait += clock64() - t0;
                    }
                    if (++idle > 20000000u) {
                        wd_code = 1;
                        break;
                    }
                    ORZ_POLL();
                    continue;
                }
                idle = 0;

                if (nst > 0) {
                    __syncwarp();
                    // ---- verify, lane = step: a step is exact only if no EARLIER step of this super-window touched what it read — its own
                    // or its lazy neighbours' candidate lists (same key) or its words entry — and no two steps write the same words
                    // entry.  The first step that fails ends the super-window and becomes the first step of the next one.
                    const bool live = (uint32_t)lane < nst;
                    uint4 e0 = make_uint4(0, 0, 0, 0), e1 = make_uint4(0, 0xffffffffu, 0, 0);
                    if (live) {
                        const uint32_t *ent = shl.ent_all[shl.sidx[lane]];
                        e0 = *reinterpret_cast<const uint4 *>(ent);
                        e1 = *reinterpret_cast<const uint4 *>(ent + 4);
#if !ORZ_V7_VERIFY_TABLE
                        shl.vkw[lane] = make_uint2(e1.x, e1.y);
#endif
                    }
                    __syncwarp();
                    const uint32_t x = e0.x, ty = e0.y & 3u, lzn = (e0.y >> 2) & 1u, room5 = (e0.y >> 3) & 31u, adv = (e0.y >> 8) & 0xffu, mlen = (e0.y >> 16) & 0xffu, wexp = e0.y >> 24;
                    const uint32_t key = e1.x, key1 = e0.z & 0x1fffffu, key2 = e0.w & 0x1fffffu, nbroom = e0.w >> 21, wu = e1.y, h2 = e1.z & 0x7fffu, lwe = e1.z >> 15, word = e1.w;
                    const uint32_t b2lo = e0.z >> 21;
                    // earlier steps of the super-window that insert into this step's bucket: more than `room` of them and the oldest
                    // candidate of its record has left the ring (matcher.rs:82-87) by the time the step is taken
                    const unsigned mb = __match_any_sync(kFull, live ? key >> 13 : 256u + (uint32_t)lane);
                    bool viol = (uint32_t)__popc(mb & lt_mask) > room5 + ORZ_V7_MUT_TIGHT || (uint32_t)lane > nbroom + ORZ_V7_MUT_TIGHT;
#if ORZ_V7_VERIFY_TABLE
                    // Constant work per step instead of one pass per earlier step.  Same key as an earlier step, and the later
                    // writers of this step's words entry: one MATCH each.  The cross questions — did an earlier step insert a node
                    // with the key of x+1 / x+2 (lazy neighbours), did one write the words entry this step read — go through a small
                    // table of lane masks (entry = the steps whose key / words index hashes there; every candidate is confirmed).
                    const bool writer = live && wu != 0xffffffffu;
                    const uint32_t widx = wu & 0x7fffu;
                    const unsigned mk = __match_any_sync(kFull, live ? key : 0x40000000u | (uint32_t)lane);
                    const unsigned mw = __match_any_sync(kFull, writer ? widx : 0x40000000u | (uint32_t)lane);
                    viol = viol || (mk & lt_mask) != 0u;
                    const unsigned wlater = mw & ~lt_mask & ~(1u << lane); // later steps that write the same words entry (the last writer wins)
                    if (live) atomicOr(&shl.vtab[vhash7(key)], 1u << lane);
                    if (writer) atomicOr(&shl.vtab[256u + vhash7(widx)], 1u << lane);
                    __syncwarp();
                    unsigned m1 = 0, m2 = 0, m3 = 0;
                    if (live) {
                        if (lzn) {
                            m1 = shl.vtab[vhash7(key1)] & lt_mask;
                            m2 = shl.vtab[vhash7(key2)] & lt_mask;
                        }
                        m3 = shl.vtab[256u + vhash7(h2)] & lt_mask;
                    }
                    __syncwarp();
                    if (live) shl.vtab[vhash7(key)] = 0u;
                    if (writer) shl.vtab[256u + vhash7(widx)] = 0u;
                    while (__ballot_sync(kFull, (m1 | m2) != 0u)) { // earlier steps whose key shares a table entry with x+1's / x+2's
                        const int b1 = m1 ? 31 - __clz(m1) : 0, b2 = m2 ? 31 - __clz(m2) : 0;
                        const uint32_t k1 = __shfl_sync(kFull, key, b1), k2 = __shfl_sync(kFull, key, b2);
                        viol = viol || (m1 && k1 == key1) || (m2 && k2 == key2);
                        if (m1) m1 &= ~(1u << b1);
                        if (m2) m2 &= ~(1u << b2);
                    }
                    while (__ballot_sync(kFull, m3 != 0u)) { // the latest earlier writer of the words entry this step read decides (lz.rs:132,208,233)
                        const int b3 = m3 ? 31 - __clz(m3) : 0;
                        const uint32_t w3 = __shfl_sync(kFull, wu, b3);
                        if (m3) {
                            if ((w3 & 0x7fffu) == h2) {
                                viol = viol || (w3 >> 15) != lwe;
                                m3 = 0u;
                            } else m3 &= ~(1u << b3);
                        }
                    }
#else
                    uint32_t wnext = 32; // the next step of the super-window that writes the same words entry (the last writer wins)
                    // (four steps per trip, their broadcast loads issued together: a lone warp pays the full shared-memory latency otherwise)
                    for (uint32_t j0 = 0; j0 < nst; j0 += 4) {
                        uint2 kw4[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) kw4[u] = shl.vkw[(j0 + (uint32_t)u) & 31u];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const uint32_t j = j0 + (uint32_t)u;
                            const uint32_t kj = kw4[u].x, wj = kw4[u].y;
                            const uint32_t wi = wj & 0x7fffu;
                            // an earlier step put a node on one of the lists this step read, or changed the words entry it read
                            const bool hit = kj == key || (lzn && (kj == key1 || kj == key2)) || (wj != 0xffffffffu && wi == h2 && (wj >> 15) != lwe);
                            viol = viol || (j < (uint32_t)lane && j < nst && hit);
                            if (j > (uint32_t)lane && j < nst && wnext == 32u && wj != 0xffffffffu && wu != 0xffffffffu && wi == (wu & 0x7fffu)) wnext = j;
                        }
                    }
                    const unsigned wlater = wnext < 32u ? 1u << wnext : 0u;
#endif
                    const unsigned vm = __ballot_sync(kFull, viol && live);
                    uint32_t nok = nst;
                    if (vm) {
                        nok = (uint32_t)(__ffs(vm) - 1);
                        base = __shfl_sync(kFull, x, (int)nok); // the failed step starts the next super-window
                        stop = 0;
                        if (STATS) n_cut++;
                    }
                    // ---- hand the steps to the committers and fold them into the shared tables
                    const bool onpath = (uint32_t)lane < nok;
                    if (qhead + nok > qlimit) { // queue space: both committers must have passed the entries about to be overwritten
                        for (;;) {
                            const uint32_t d1 = bcast0(ld_vol_u32(v_qdone)), d2 = bcast0(ld_vol_u32(v_qitem));
                            qlimit = (d1 < d2 ? d1 : d2) + (uint32_t)kQueue7 - 8u;
                            if (qhead + nok <= qlimit) break;
                            ORZ_POLL();
                        }
                    }
                    const uint32_t pty = __shfl_up_sync(kFull, ty, 1);
                    const uint32_t my_al = lane == 0 ? after_literal : (pty == kTyLit ? 1u : 0u); // lz.rs:229
                    if (onpath) {
                        uint4 rec;
                        rec.x = x | ty << 25 | my_al << 27;
                        rec.y = word;                                   // ordinal of the matched node; the committers number the inserts
                        rec.z = key | (ty == kTyMatch ? mlen : 0u) << 21;
                        rec.w = wexp | (lwe & 0xffu) << 8 | b2lo << 16;
                        shl.queue[(qhead + (uint32_t)lane) % kQueue7] = rec;
                        if ((STATS || ORZ_V7_TRACE_ALWAYS) && B.dbg_pos && x - B.dbg_pos < 3u) {
                            const uint32_t tv[9] = {x, ty, mlen, word, (uint32_t)lane, nst, nok, key, wexp};
                            trace7(S, 3u, tv, 9);
                        }
                        atomicAdd(&shl.bcnt[key >> 13], 1u);
                        atomicMax(&shl.lastkey[(key ^ (key >> 9)) & 4095u], x);
                        if (ty != kTyWord && !(wlater & ((nok < 32u ? 1u << nok : 0u) - 1u))) shl.words[wu & 0x7fffu] = (uint16_t)(wu >> 15);
                    }
                    after_literal = __shfl_sync(kFull, ty, (int)nok - 1) == kTyLit ? 1u : 0u;
                    (void)adv;
                    qhead += nok;
                    n_items += nok;
                    spos = base;
                    __syncwarp();
                    if (lane == 0) {
                        *v_qhead = qhead;
                        *v_spos = spos;
                        *reinterpret_cast<volatile uint32_t *>(&shl.lead_pos) = spos;
                    }
                    if (STATS) {
                        n_handed += nok;
                        lag_steps += qhead - ld_vol_u32(v_qdone);
                        lag_pos += spos - ld_vol_u32(&shl.fpos);
                    }
                }
                const long long t3 = STATS ? clock64() : 0;
                if (STATS) {
                    cyc_commit += t3 - t1;
                    n_stopc[stop < 5 ? stop : 0]++;
                }
                if (stop == 1 && spos < sbuf_len && n_items < items_cap) {
                    // ---- REPAIR: the record of spos (or of a lazy neighbour it needs) is stale.  Re-resolve it in place from the step
                    // log (repair_record7); a record that cannot be repaired that way is rebuilt from HBM after the committer drained.
                    if (STATS) n_repair++;
                    ByteSrc7 src;
                    src.win = B.win;
                    src.ring = shl.inring;
                    const uint32_t flo = shl.fed_lo0, fhi = bcast0(ld_vol_u32(&shl.fed_hi));
                    src.lo = spos - 2048u > flo ? spos - 2048u : flo;
                    src.hi = fhi;
                    bool lzneed0 = false;
                    for (uint32_t k = 0; k < 3; k++) {
                        const uint32_t y = spos + k;
                        if (y > last_pos) break;
                        // the records the decide team found stale (one that became stale through this very hand-over shows up in the next round)
                        const uint32_t yslot = slot7<NBW>(y);
                        uint32_t *ys = shl.sum[yslot];
                        const uint32_t have_y = y;
                        const bool stale = (rmask >> k) & 1u;
                        if (have_y == y && stale && (k == 0 || lzneed0)) {
                            { // the record may be waiting for, or in the middle of, a refresh by its builder: take the request back; a
                              // refresh that is already running is waited for, then the round is simply taken again
                                const uint32_t w = yslot >> 5, bit = 1u << (yslot & 31u);
                                if ((bcast0(ld_vol_u32(&shl.req[w])) | bcast0(ld_vol_u32(&shl.busy[w]))) & bit) {
                                    if (lane == 0) atomicAnd(&shl.req[w], ~bit);
                                    __threadfence();
                                    __syncwarp();
                                    if (bcast0(ld_vol_u32(&shl.busy[w])) & bit) {
                                        if (STATS) n_rwait++;
                                        while (bcast0(ld_vol_u32(&shl.busy[w])) & bit) ORZ_POLL();
                                        break;
                                    }
                                }
                            }
#if ORZ_V7_PROV
                            uint32_t rdbg[4] = {0, 0, 0, 0};
                            const int how = repair_record7(shl, B, src, y, yslot, spos, qhead, lane, rdbg);
#else
                            const int how = repair_record7(shl, B, src, y, yslot, spos, qhead, lane);
#endif
                            if (how && lane == 0) ys[10] = sumcheck7(y, ys[1], ys[2], ys[3], ys[4], ys[5], ys[6], ys[7], ys[8], ys[9]); // (the repair rewrote summary words)
                            __syncwarp();
                            if (STATS) n_how[how & 3]++;
#if ORZ_V7_PROV
                            if (lane == 0 && S->prov) {
                                const size_t pi = (size_t)B.prov_base + (y - kPrematch);
                                if (y >= kPrematch && pi < (size_t)S->prov_cap) {
                                    const uint32_t *pex = shl.ext[yslot];
                                    S->prov[4 * pi + 2] = make_uint4(spos, (uint32_t)how | 0x100u | (rmask & 7u) << 12 | k << 16, ys[3], shl.rc_pos[0]);
                                    S->prov[4 * pi + 3] = make_uint4(rdbg[0], rdbg[1], rdbg[2], rdbg[3]);
                                    (void)pex;
                                }
                            }
#endif
                            if ((STATS || ORZ_V7_TRACE_ALWAYS) && B.dbg_pos && y - B.dbg_pos < 3u && lane == 0) {
                                const uint32_t *ex = shl.ext[yslot];
                                const uint32_t tv[24] = {y, spos, (uint32_t)how, ys[3], ys[4], ys[5], ys[6], shl.rc_pos[0], shl.rc_exp[0], shl.rc_ord[0], ex[0], ex[1], ys[1],
                                                         shl.bcnt[(ys[1] >> 13) & 0xffu], yslot, qhead, ex[2], ex[3], ex[4], ex[5], ex[6], ex[7], ex[8], ex[9]};
                                trace7(S, 2u, tv, 24);
                            }
                            if (!how) {
                                if (STATS) n_rep_hard++;
                                while (bcast0(ld_vol_u32(v_qdone)) != qhead) ORZ_POLL();
                                __threadfence();
                                build_record7(B, y, spos, shl.bcnt, ys, shl.ext[yslot], yslot, lane, 2);
                            }
                            if (STATS) n_rep_rec++;
                        }
                        __syncwarp();
                        if (k == 0) {
                            const uint32_t m0 = have_y == y ? (ys[3] & 0xffu) : 0u;
                            lzneed0 = m0 > 0 && m0 < (uint32_t)(kMatchMax / 2);
                        }
                    }
                    __syncwarp();
                    if (STATS) cyc_repair += clock64() - t3;
                }
            }
            __syncwarp();
            if (lane == 0) *v_done = 1;
            __threadfence_block();
            __syncwarp();
            bar_arrive7(1, kTeamThreads); // the team wakes up, sees `done` and leaves
            if (lane == 0) {
                if (wd_code) {
                    S->status = -7;
                    S->dbg[6] = (unsigned long long)wd_code;
                    S->dbg[7] = spos;
                }
                S->spos = spos;
                S->after_literal = after_literal;
                if (!wd_code) S->dbg[7] += ld_vol_u32(&shl.torn);
                (buf ? S->n_items_alt : S->n_items) = n_items;
                if (STATS) {
                    S->dbg[0] += n_items;
                    S->dbg[1] += n_windows;
                    S->dbg[2] += n_wait;
                    S->dbg[3] += n_repair;
                    S->dbg7[0] += (unsigned long long)(clock64() - t_begin);
                    S->dbg7[1] += (unsigned long long)cyc_load;
                    S->dbg7[2] += (unsigned long long)cyc_walk;
                    S->dbg7[3] += (unsigned long long)cyc_commit;
                    S->dbg7[4] += (unsigned long long)cyc_repair;
                    S->dbg7[5] += (unsigned long long)cyc_wait;
                    S->dbg7[6] += n_rep_rec;
                    S->dbg7[7] += n_rep_hard;
                    S->dbg7[9] += lag_steps;
                    S->dbg7[10] += lag_pos;
                    S->dbg7[11] += n_cut;
                    S->dbg7[12] += probe_alu;
                    S->dbg7[13] += probe_shfl;
                    S->dbg7[14] += probe_lds;
                    S->dbg7[15] += probe_n;
                    for (int i = 0; i < 5; i++) S->dbg7[16 + i] += n_stopc[i];
                    for (int i = 1; i < 4; i++) S->dbg7[20 + i] += n_how[i];
                    S->dbg[6] += n_rwait;
                }
            }
        } else if (rank == 0 && warp == 1) {
            // ============================================================== list committer: key-list appends, then the frontier.
            // This is all the builders read, so the frontier trails the parse by one HBM round trip per 32 steps.
            volatile uint32_t *v_qhead = &shl.qhead;
            volatile uint32_t *v_qdone = &shl.qdone;
            uint2 *klist = S->klist;
            uint32_t *kcnt = S->kcnt;
            uint32_t tail = 0;
            const unsigned lt_mask = (1u << lane) - 1u;
            for (;;) {
                const uint32_t dn = bcast0(ld_vol_u32(v_done));
                const uint32_t head = bcast0(ld_vol_u32(v_qhead));
                if (head == tail) {
                    if (dn) break;
#ifdef ORZ_SIMT_EMU
                    ORZ_POLL();
#else
                    __nanosleep(160);
#endif
                    continue;
                }
                const uint32_t n = head - tail < 32u ? head - tail : 32u;
                if (lane == 0) V7_TRACE("C tail %u n %u\n", tail, n);
                const bool act = (uint32_t)lane < n;
                uint4 rec = make_uint4(0, 0, 0, 0);
                if (act) rec = shl.queue[(tail + (uint32_t)lane) % kQueue7];
                const uint32_t pos = rec.x & 0x1ffffffu, ty = (rec.x >> 25) & 3u;
                const uint32_t key = rec.z & 0x1fffffu, mlen = rec.z >> 21, c = key >> 13;
                const uint32_t kidx = c * kHeadsStride + (key & 0x1fffu);
                uint32_t kc = 0;
                if (act) kc = ldcg_u32(&kcnt[kidx]);
                ORZ_COST(450);
                // the bucket's insert ordinal of every step (matcher.rs:63: the ring head advances by one per update)
                const unsigned same_c = __match_any_sync(kFull, act ? c : (0x10000u | (uint32_t)lane));
                const uint32_t cb = act ? shl.lcnt[c] : 0u;
                __syncwarp();
                const uint32_t ord = (cb + (uint32_t)__popc(same_c & lt_mask)) & kOrdMask;
                if (act && !(same_c >> lane >> 1)) shl.lcnt[c] = cb + (uint32_t)__popc(same_c);
                // append in step order (matcher.rs:115-121: the new node becomes the head of its hash entry's chain)
                const unsigned same_k = __match_any_sync(kFull, act ? kidx : (0x80000000u | (uint32_t)lane));
                if (act) {
                    const uint32_t rk = (uint32_t)__popc(same_k & lt_mask);
                    klist[(size_t)kidx * kListCap + (kc + rk) % kListCap] = kentry_make(pos, ty == kTyMatch ? mlen : 0u, ord);
                }
                __threadfence();
                __syncwarp();
                if (act && !(same_k & ~lt_mask & ~(1u << lane))) kcnt[kidx] = kc + (uint32_t)__popc(same_k);
                __threadfence();
                __syncwarp();
                tail += n;
                const uint32_t lastpos = __shfl_sync(kFull, pos, (int)n - 1);
                if (lane == 0) {
                    st_release_u32(&shl.fpos, lastpos + 1);
                    *v_qdone = tail;
                }
                __syncwarp();
            }
        } else if (rank == 0 && warp == 3) {
            // ============================================================== item committer: lenmin bytes and items.  Nothing on
            // the parse or build path reads what it writes (match_len_min only shapes the emitted length code), so it runs at
            // its own pace behind the queue.
            volatile uint32_t *v_qhead = &shl.qhead;
            volatile uint32_t *v_qitem = &shl.qitem;
            Item *items = buf ? S->items_alt : S->items;
            uint8_t *lmarr = S->lm;
            uint32_t tail = 0;
            const unsigned lt_mask = (1u << lane) - 1u;
            for (;;) {
                const uint32_t dn = bcast0(ld_vol_u32(v_done));
                const uint32_t head = bcast0(ld_vol_u32(v_qhead));
                if (head == tail || (head - tail < 24u && !dn)) { // (batches of at least 24 steps: fewer, fuller passes)
                    if (dn && head == tail) break;
#ifdef ORZ_SIMT_EMU
                    ORZ_POLL();
#else
                    __nanosleep(600);
#endif
                    continue;
                }
                const uint32_t n = head - tail < 32u ? head - tail : 32u;
                const bool act = (uint32_t)lane < n;
                uint4 rec = make_uint4(0, 0, 0, 0);
                if (act) rec = shl.queue[(tail + (uint32_t)lane) % kQueue7];
                const uint32_t ty = (rec.x >> 25) & 3u, al = (rec.x >> 27) & 1u;
                const uint32_t word = rec.y, key = rec.z & 0x1fffffu, mlen = rec.z >> 21;
                const uint32_t c = key >> 13;
                const uint32_t wexp = rec.w & 0xffu, unl = (rec.w >> 8) & 0xffu, byte = (rec.w >> 16) & 0xffu;
                const unsigned same_c = __match_any_sync(kFull, act ? c : (0x10000u | (uint32_t)lane));
                const uint32_t cb = act ? shl.icnt[c] : 0u;
                __syncwarp();
                const uint32_t ord = (cb + (uint32_t)__popc(same_c & lt_mask)) & kOrdMask;
                if (act && !(same_c >> lane >> 1)) shl.icnt[c] = cb + (uint32_t)__popc(same_c);
                const uint32_t ro = (ord - 1 - word) & 0xfffu; // reduced offset (matcher.rs:185): inserts since the matched node
                const bool is_m = act && ty == kTyMatch;
                // ring slots: the insert goes to (head + 1) % 4094 where head = slot of the previous insert (matcher.rs:63,77);
                // bucket insert number `ord` (0-based) therefore lands in slot (ord + 1) % 4094
                const uint32_t slot = (ord + 1) % kRing;
                const uint32_t mord = (ord - 1 - ro) & kOrdMask;
                const uint32_t mslot = (mord + 1) % kRing;
                const uint32_t wkey_ins = act ? (c << 12 | slot) : (0x80000000u | (uint32_t)lane);
                const uint32_t wkey_m = is_m ? (c << 12 | mslot) : (0x40000000u | (uint32_t)lane);
                // lenmin of the matched node as stored; the pass below applies the effects of earlier steps of this batch
                uint32_t cur_lm = 0;
                if (is_m) cur_lm = ldcg_u8(&lmarr[c * kNodesStride + mslot]);
                ORZ_COST(450);
                // matcher.rs:62-80 in step order: [bump the matched node's lenmin] then [write the new node, lenmin 0].
                // Only steps that can interact need the ordered pass: two steps matching the same node, or a step matching a node
                // inserted by an earlier step of this batch (its reduced offset is then < 32).
                const unsigned grp = __match_any_sync(kFull, wkey_m);
                const unsigned involved = __ballot_sync(kFull, is_m && (ro < 32u || (grp & ~(1u << lane)) != 0u));
                uint32_t new_lm = 0;     // value this lane's bump writes
                bool bump = false;
                bool ins_dead = false, bump_dead = false; // a later step of the batch overwrites the location
                const int last = involved ? 31 - __clz(involved) : -1; // effects only flow INTO involved steps: stop there
                for (int i = 0; i <= last; i++) {
                    const uint32_t wm_i = __shfl_sync(kFull, wkey_m, i);
                    const uint32_t wi_i = __shfl_sync(kFull, wkey_ins, i);
                    uint32_t bv = 0xffffffffu;
                    if (lane == i && is_m && cur_lm <= mlen) {
                        bump = true;
                        new_lm = mlen + 1 < 127u ? mlen + 1 : 127u;
                        bv = new_lm;
                    }
                    bv = __shfl_sync(kFull, bv, i);
                    if (lane > i) {
                        if (bv != 0xffffffffu && wkey_m == wm_i) cur_lm = bv; // an earlier step bumped the node this lane matched
                        if (wkey_m == wi_i) cur_lm = 0;                        // ... or it IS that step's new node
                    }
                    if (lane < i) { // later writes to the same byte make this lane's stores dead
                        if (wkey_ins == wi_i || (bv != 0xffffffffu && wkey_ins == wm_i)) ins_dead = true;
                        if (bump && (wkey_m == wi_i || (bv != 0xffffffffu && wkey_m == wm_i))) bump_dead = true;
                    }
                }
                if (is_m && lane > last && cur_lm <= mlen) { // no other step of the batch touches this node: plain matcher.rs:66-71
                    bump = true;
                    new_lm = mlen + 1 < 127u ? mlen + 1 : 127u;
                }
                if (bump && wkey_m == wkey_ins) bump_dead = true; // matcher.rs:66-77: the step's own insert overwrites the bumped slot
                // ---- items (lz.rs:173-236)
                const uint32_t ctx = c | al << 8;
                const unsigned same_ctx = __match_any_sync(kFull, act ? ctx : (0x10000u | (uint32_t)lane));
                const uint32_t cbase = act ? shl.cnt[ctx] : 0u;
                __syncwarp();
                if (act) {
                    Item it;
                    it.ctx = (uint16_t)ctx;
                    it.unlikely = (uint8_t)unl;
                    it.pad = 0;
                    it.seq = cbase + (uint32_t)__popc(same_ctx & lt_mask);
                    if (ty == kTyMatch) {
                        uint32_t roid, robitlen, robits;
                        roid_encode(ro, roid, robitlen, robits);                    // lz.rs:148
                        const uint32_t mexp = wexp > (uint32_t)kMatchMin ? wexp : (uint32_t)kMatchMin; // matcher.rs:187-188
                        const uint32_t mmin = cur_lm > (uint32_t)kMatchMin ? cur_lm : (uint32_t)kMatchMin;
                        uint32_t eml;                                               // lz.rs:173-177
                        if (mlen > mexp) eml = mlen - mmin;
                        else if (mlen < mexp) eml = mlen - mmin + 1;
                        else eml = 0;
                        eml &= 0xffu;
                        const uint32_t lenid = eml < (uint32_t)(kLenidSize - 1) ? eml : (uint32_t)(kLenidSize - 1);
                        it.symbol = (uint16_t)(256 + roid * kLenidSize + lenid);
                        it.robits = (uint16_t)robits;
                        it.robitlen = (uint8_t)robitlen;
                        it.eml = (uint8_t)eml;
                        it.flags = (uint8_t)(al | 2u);
                    } else {
                        it.symbol = ty == kTyWord ? (uint16_t)kWordSym : (uint16_t)byte;
                        it.robits = 0;
                        it.robitlen = 0;
                        it.eml = 0;
                        it.flags = (uint8_t)al;
                    }
                    items[tail + (uint32_t)lane] = it;
                    if (!(same_ctx & ~lt_mask & ~(1u << lane))) shl.cnt[ctx] = cbase + (uint32_t)__popc(same_ctx); // group's last lane
                    // ---- ring bytes
                    if (bump && !bump_dead) lmarr[c * kNodesStride + mslot] = (uint8_t)new_lm;
                }
                // An earlier step may bump the node whose ring slot a later step of the batch re-uses (a match 4087 inserts back, found by
                // the randomized emulator soak): the new node's lenmin 0 must land after that bump (matcher.rs:66-77 in step order).
                // Two stores from different lanes are ordered by the barrier, not by the warp happening to run in lock step.
                __syncwarp();
                if (act && !ins_dead) lmarr[c * kNodesStride + slot] = 0;
                __threadfence(); // the next batch's lenmin loads (other lanes, L2) must see these stores
                __syncwarp();
                tail += n;
                if (lane == 0) *v_qitem = tail;
                __syncwarp();
            }
        } else if (rank == 0 && warp == 2) {
            // ============================================================== feeder: window bytes around the parse position
            volatile uint32_t *v_fed = &shl.fed_hi;
            uint32_t fed = shl.fed_lo0;
            const uint32_t top = S->win_hi & ~15u;
            for (;;) {
                if (bcast0(ld_vol_u32(v_done))) break;
                const uint32_t sp = bcast0(ld_vol_u32(v_spos));
                uint32_t target = (sp + kInAhead) & ~15u;
                if (target > top) target = top;
                if (fed >= target) {
#ifdef ORZ_SIMT_EMU
                    ORZ_POLL();
#else
                    __nanosleep(800);
#endif
                    continue;
                }
                while (fed < target) {
                    const uint32_t p = fed + 16u * (uint32_t)lane;
                    if (p < target) *reinterpret_cast<uint4 *>(&shl.inring[(p & (kInRing - 1)) >> 2]) = *reinterpret_cast<const uint4 *>(B.win + p);
                    fed = fed + 512u < target ? fed + 512u : target;
                }
                __threadfence_block();
                __syncwarp();
                if (lane == 0) *v_fed = fed;
            }
        } else if (rank == 0 && warp >= 4 && warp < 4 + kTeam7) {
            // ============================================================== decide team: warp 4 + t takes block t of every round
            const uint32_t t = (uint32_t)(warp - 4);
            for (;;) {
                bar_sync7(1, kTeamThreads); // parked until the resolver starts a round (or ends the launch)
                if (bcast0(ld_vol_u32(v_done))) break;
                const uint32_t tb = bcast0(ld_vol_u32(&shl.team_base));
                decide_block7<NBW>(shl, tb, tb + 30u * t, t == 0, sbuf_len, lane, shl.pk_all[t], &shl.ent_all[t * 32u], S, (STATS || ORZ_V7_TRACE_ALWAYS) ? B.dbg_pos : 0u);
                __threadfence_block();
                __syncwarp();
                bar_arrive7(2, kTeamThreads);
            }
        } else if (rank != 0 || (K == 1 && warp >= 4 + kTeam7 && (warp & 3) != 0)) {
            // ============================================================== builders
            int bi;
            if (rank == 0) bi = warp - 4 - kTeam7 - (warp - 4 - kTeam7 + 3) / 4;
            else bi = NB0 + ((int)rank - 1) * NW + warp;
            unsigned long long nbuild = 0, nrefresh = 0;
            long long cyc_build = 0;
            const uint32_t last_pos = sbuf_len + 1;
            uint32_t y = S->spos;
            y += ((uint32_t)bi + (uint32_t)NBW - y % (uint32_t)NBW) % (uint32_t)NBW; // first position >= spos owned by this builder
            const uint32_t first_own = y;
            for (;;) {
                const uint4 pub = ld_volatile_v4_generic(&sh.spos); // spos, lead_pos, fpos, done
                if (bcast0(pub.w)) break;
                const uint32_t sp = bcast0(pub.x), lp = bcast0(pub.y), F = bcast0(pub.z);
                if (y < sp) y += ((sp - y + (uint32_t)NBW - 1) / (uint32_t)NBW) * (uint32_t)NBW; // the parse went past for good: skip
                // ---- refresh requests for the records this builder owns that the parse has not reached yet (its newest two positions)
                bool refreshed = false;
                uint32_t rw_idx = 0xffffffffu, rw = 0;
                for (uint32_t j = 1; j <= 2u && j <= (uint32_t)kRing7 / (uint32_t)NBW; j++) {
                    if (y < first_own + j * (uint32_t)NBW) break;
                    const uint32_t yy = y - j * (uint32_t)NBW;
                    if (yy < sp || yy > last_pos) continue;
                    const uint32_t rslot = slot7<NBW>(yy), w = rslot >> 5, bit = 1u << (rslot & 31u);
                    if (w != rw_idx) {
                        rw = bcast0(ld_vol_u32(&sh.req[w]));
                        rw_idx = w;
                    }
                    if (!(rw & bit)) continue;
                    // take the request: `busy` goes up before `req` comes down, so whoever finds the request gone sees the rebuild running
                    uint32_t old = 0;
                    if (lane == 0) {
                        atomicOr(&sh.busy[w], bit);
                        __threadfence();
                        old = atomicAnd(&sh.req[w], ~bit);
                    }
                    old = bcast0(old);
                    if (old & bit) {
                        if (lane == 0) V7_TRACE("B%d refresh y %u F %u sp %u\n", bi, yy, F, sp);
                        build_record7(B, yy, F, sh.bcnt, sh.sum[rslot], sh.ext[rslot], rslot, lane, 1);
                        if (STATS) nrefresh++;
                        refreshed = true;
                    }
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) atomicAnd(&sh.busy[w], ~bit);
                    rw_idx = 0xffffffffu;
                }
                if (refreshed) continue; // (the frontier and the parse position have moved on)
                // (kLeadMax from `spos` keeps the slot of y - NBW * SL behind everything a fall-back could still need)
                if (y > last_pos || y > (lp > sp ? lp : sp) + kLead || y > sp + kLeadMax) {
#ifdef ORZ_SIMT_EMU
                    ORZ_POLL();
#else
                    __nanosleep(400);
#endif
                    continue;
                }
                if (lane == 0) V7_TRACE("B%d y %u F %u sp %u\n", bi, y, F, sp);
                const uint32_t yslot = slot7<NBW>(y);
                { // a request left behind by the slot's previous position must not hit this one
                    const uint32_t w = yslot >> 5, bit = 1u << (yslot & 31u);
                    if (w != rw_idx) rw = bcast0(ld_vol_u32(&sh.req[w]));
                    if ((rw & bit) && lane == 0) atomicAnd(&sh.req[w], ~bit);
                }
                const long long tb = STATS ? clock64() : 0;
                // (the list loads — L2, ld.cg — are issued after the frontier was read: their address depends on it)
                build_record7(B, y, F, sh.bcnt, sh.sum[yslot], sh.ext[yslot], yslot, lane);
                if (STATS) {
                    nbuild++;
                    cyc_build += clock64() - tb;
                }
                y += (uint32_t)NBW;
            }
            if (STATS && lane == 0) {
                atomicAdd(&S->dbg[4], nbuild);
                atomicAdd(&S->dbg7[8], (unsigned long long)cyc_build);
                atomicAdd(&S->dbg[5], nrefresh);
            }
        }
    }
    __syncthreads();
    if (K > 1) cluster_barrier(); // no remote builder touches CTA 0's shared memory after this point
    if (active && rank == 0) {
        for (int i = tid; i < 32768 / 8; i += NT) reinterpret_cast<uint4 *>(S->words)[i] = reinterpret_cast<const uint4 *>(shl.words)[i];
        for (int i = tid; i < kNumCtx; i += NT) (buf ? S->ctx_count_alt : S->ctx_count)[i] = shl.cnt[i];
        for (int i = tid; i < kNumBuckets; i += NT) S->bcnt[i] = shl.bcnt[i];
    }
}

// lz.rs:82-87 for the key lists: positions are rebased, entries whose node died (matcher.rs:82-87,123-133: pos <= forward_len)
// become pos 0, which ends a candidate list there.
__device__ __forceinline__ void enc7_forward_rebase(uint2 *klist, const uint32_t *kcnt, uint32_t forward_len, size_t gid, size_t gsz) {
    const size_t total = (size_t)kNumBuckets * kHeadsStride * kListCap;
    for (size_t i = gid; i < total; i += gsz) {
        if (kcnt[i / kListCap] == 0) continue;
        uint2 e = klist[i];
        const uint32_t p = kentry_pos(e);
        const uint32_t np = p > forward_len ? p - forward_len : 0u;
        e.x = (e.x & 0xfe000000u) | np;
        klist[i] = e;
    }
}

} // namespace orz
