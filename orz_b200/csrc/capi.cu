// capi.cu — host side of the C ABI declared in include/orz_b200.h.
//
// Owns device memory (one arena per batch), drives the kernels, mirrors the reference's stream
// driver (lib.rs:58-129) and operator interface (lz.rs:75-95, 353-366).  No CPU codec lives here:
// if CUDA is unavailable every entry point fails with ORZ_ECUDA.
#include "../../include/orz_b200.h"
#include "orz_dev.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace orz {
void launch_parse(EncStream *, int, bool v7, int buf, cudaStream_t);
void launch_emit(EncStream *, int, int buf, cudaStream_t);
void launch_chunk_done(EncStream *, int, int buf, cudaStream_t);
void launch_enc_begin(EncStream *, int, cudaStream_t);
void launch_enc_advance(EncStream *, int, cudaStream_t);
void launch_enc_forward(EncStream *, uint32_t, bool v7, cudaStream_t);
bool parse7_stats_enabled();
void launch_decode(DecStream *, int, cudaStream_t);
void launch_dec_forward(DecStream *, uint32_t, cudaStream_t);
} // namespace orz

using namespace orz;

#define CU_TRY(expr)                                                                                        \
    do {                                                                                                    \
        cudaError_t e_ = (expr);                                                                            \
        if (e_ != cudaSuccess) {                                                                            \
            if (getenv("ORZ_B200_DEBUG")) fprintf(stderr, "orz_b200: %s failed: %s\n", #expr, cudaGetErrorString(e_)); \
            return e_ == cudaErrorMemoryAllocation ? ORZ_ENOMEM : ORZ_ECUDA;                                 \
        }                                                                                                   \
    } while (0)

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Every entry point works on the batch's device and leaves the caller's current device as it found it.
struct DevGuard {
    int prev = -1;
    bool ok = false;
    explicit DevGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(device) == cudaSuccess;
    }
    ~DevGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
#define DEV_GUARD(dev)              \
    DevGuard dev_guard_(dev);       \
    if (!dev_guard_.ok) return ORZ_ECUDA

// stage A v7 (parse_v7.cuh) keeps ~0.6 GB of key lists per stream: used for up to this many streams per batch
static const int kV7MaxStreams = 8;
static const size_t kV7MaxDepth = 48;

#ifdef ORZ_SIMT_EMU
// Only the test-tier emulator build (tests/simt/) exports this marker; orz_b200/_lib.py refuses to load a library
// that has it unless the test suite asks for it explicitly, so the emulator can never stand in for the GPU.
extern "C" int orz_b200_is_emulator(void) { return 1; }
#endif

// ------------------------------------------------------------------------------------------------
// timing: CUDA events around every launch, on the launching stream
// ------------------------------------------------------------------------------------------------
struct KernelTimer {
    struct Rec {
        cudaEvent_t a, b;
        int cls;
    };
    std::vector<Rec> pool;
    size_t used = 0;
    float ms[5] = {0, 0, 0, 0, 0};
    int launches[5] = {0, 0, 0, 0, 0};
    void reset() {
        used = 0;
        for (int i = 0; i < 5; i++) ms[i] = 0, launches[i] = 0;
    }
    int begin(int cls, cudaStream_t st) {
        if (used == pool.size()) {
            Rec r;
            cudaEventCreate(&r.a);
            cudaEventCreate(&r.b);
            r.cls = cls;
            pool.push_back(r);
        }
        pool[used].cls = cls;
        cudaEventRecord(pool[used].a, st);
        return (int)used++;
    }
    void end(int id, cudaStream_t st) { cudaEventRecord(pool[id].b, st); }
    void collect() { // after a stream synchronize
        for (size_t i = 0; i < used; i++) {
            float t = 0;
            if (cudaEventElapsedTime(&t, pool[i].a, pool[i].b) == cudaSuccess) {
                ms[pool[i].cls] += t;
                launches[pool[i].cls]++;
            }
        }
    }
    ~KernelTimer() {
        for (auto &r : pool) {
            cudaEventDestroy(r.a);
            cudaEventDestroy(r.b);
        }
    }
};

// ------------------------------------------------------------------------------------------------
// batch object
// ------------------------------------------------------------------------------------------------
struct orz_batch {
    int device = 0;
    int mode = 0; // 0 encode, 1 decode
    int n = 0;
    orz_lzcfg cfg{};
    cudaStream_t st = nullptr;
    cudaStream_t st2 = nullptr;      // v7 encode: stages B-D of chunk k run here while stage A of chunk k+1 runs on `st`
    cudaEvent_t ev_parse[2] = {nullptr, nullptr}, ev_emit[2] = {nullptr, nullptr};
    uint8_t *arena = nullptr;
    size_t arena_size = 0;
    // state region inside the arena that must be reset before each run: [state_off, state_off+state_size)
    std::vector<EncStream> enc;  // host mirrors (templates)
    std::vector<DecStream> dec;
    EncStream *d_enc = nullptr;
    DecStream *d_dec = nullptr;
    std::vector<uint8_t *> d_in, d_out; // per-stream device buffers
    std::vector<size_t> in_cap, in_len, out_cap;
    std::vector<uint8_t *> zero_lo;     // per-stream [zero_lo, zero_hi) zero-filled before each run
    std::vector<uint8_t *> zero_hi;
    std::vector<int16_t *> heads;       // per-stream heads (0xff filled)
    std::vector<int32_t> status;
    std::vector<uint64_t> result_len;
    int32_t *h_status_pinned = nullptr;
    bool chunk_mode = false; // operator interface: host drives chunks
    bool v7 = false;         // encoder state in the v7 layout (key lists), else ring records + hash chains
    KernelTimer timer;
    bool ran = false;
    void *prov = nullptr;    // ORZ_B200_PROV: per-step provenance of stream 0 (A/B builds, see EncStream::prov)
};

static int check_device(int device) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return ORZ_ECUDA;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return ORZ_ECUDA;
    if (p.major != 10) return ORZ_ECUDA; // the library embeds an sm_100a cubin only
    return ORZ_OK;
}

extern "C" int orz_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) return 0;
    int ok = 0;
    for (int i = 0; i < count; i++)
        if (check_device(i) == ORZ_OK) ok++;
    return ok;
}

extern "C" const char *orz_strerror(int code) {
    switch (code) {
    case ORZ_OK: return "ok";
    case ORZ_EINVAL: return "invalid data";
    case ORZ_ENOSPC: return "output buffer too small";
    case ORZ_EEOF: return "unexpected end of file";
    case ORZ_ENOMEM: return "out of memory";
    case ORZ_ECUDA: return "CUDA device/kernel unavailable (no CPU fallback)";
    case ORZ_EARG: return "invalid argument";
    case -7: return "internal: stage A watchdog tripped (probe protocol bug)";
    default: return "unknown error";
    }
}

extern "C" int orz_level_cfg(int level, orz_lzcfg *cfg) { // main.rs:97-102
    if (!cfg) return ORZ_EARG;
    switch (level) {
    case 0: *cfg = {5, 3, 2}; return ORZ_OK;
    case 1: *cfg = {15, 9, 6}; return ORZ_OK;
    case 2: *cfg = {45, 27, 18}; return ORZ_OK;
    default: return ORZ_EARG;
    }
}

extern "C" size_t orz_encode_bound(size_t n) { return n + n / 2 + (n / 65536 + 2) * 4096 + 65536; }

struct ArenaPlan {
    size_t off = 0;
    size_t take(size_t bytes, size_t align = 256) {
        off = align_up(off, align);
        size_t r = off;
        off += bytes;
        return r;
    }
};

static void batch_destroy(orz_batch *b) {
    if (!b) return;
    DevGuard g(b->device);
    if (b->st) cudaStreamSynchronize(b->st);
    if (b->arena) cudaFree(b->arena);
    if (b->prov) cudaFree(b->prov);
    if (b->h_status_pinned) cudaFreeHost(b->h_status_pinned);
    if (b->st2) {
        cudaStreamSynchronize(b->st2);
        cudaStreamDestroy(b->st2);
    }
    for (int i = 0; i < 2; i++) {
        if (b->ev_parse[i]) cudaEventDestroy(b->ev_parse[i]);
        if (b->ev_emit[i]) cudaEventDestroy(b->ev_emit[i]);
    }
    if (b->st) cudaStreamDestroy(b->st);
    delete b;
}

// full_window: allocate the whole 32 MiB window even for short inputs (operator interface)
static orz_batch *batch_create(int device, int mode, const orz_lzcfg *cfg, int n, const size_t *in_len, const size_t *out_cap,
                               bool chunk_mode, int *err) {
    *err = ORZ_OK;
    if (n <= 0 || (mode == 0 && !cfg)) {
        *err = ORZ_EARG;
        return nullptr;
    }
    if ((*err = check_device(device)) != ORZ_OK) return nullptr;
    DevGuard g(device);
    if (!g.ok) {
        *err = ORZ_ECUDA;
        return nullptr;
    }
    orz_batch *b = new (std::nothrow) orz_batch();
    if (!b) {
        *err = ORZ_ENOMEM;
        return nullptr;
    }
    b->device = device;
    b->mode = mode;
    b->n = n;
    b->chunk_mode = chunk_mode;
    if (cfg) b->cfg = *cfg;
    if (mode == 0) {
        const bool deep = cfg->match_depth > kV7MaxDepth || cfg->lazy_match_depth1 > kV7MaxDepth || cfg->lazy_match_depth2 > kV7MaxDepth;
        b->v7 = n <= kV7MaxStreams && (chunk_mode || !deep) && !getenv("ORZ_B200_LEGACY_STAGE_A");
    }
    const bool v7 = b->v7;
    b->d_in.resize(n);
    b->d_out.resize(n);
    b->in_cap.resize(n);
    b->in_len.assign(n, 0);
    b->out_cap.resize(n);
    b->zero_lo.resize(n);
    b->zero_hi.resize(n);
    b->heads.resize(n);
    b->status.assign(n, 0);
    b->result_len.assign(n, 0);

    ArenaPlan plan;
    struct Offs {
        size_t in, out, win, nodes, heads, words, symranks, items, items_alt, sorted, payload, zero_lo, zero_hi, klist, kcnt, lm, gext;
        size_t win_alloc;
        uint32_t win_lo, win_hi, items_cap, payload_cap;
    };
    std::vector<Offs> offs(n);
    const size_t ctl_off = plan.take(mode == 0 ? sizeof(EncStream) * n : sizeof(DecStream) * n);
    for (int i = 0; i < n; i++) {
        Offs &o = offs[i];
        const size_t len = in_len[i];
        b->in_cap[i] = len;
        b->out_cap[i] = out_cap[i];
        o.in = plan.take(len + 64);
        if (mode == 0) {
            o.out = plan.take(out_cap[i] + 64);
            const bool one_block = !chunk_mode && len <= kBlockNew;
            if (one_block) { // only positions near [PREMATCH, PREMATCH+len) are ever touched
                o.win_lo = (kPrematch - kSentinel) & ~15u;
                o.win_hi = (uint32_t)(kPrematch + len + kSentinel + kWinSlack);
            } else {
                o.win_lo = 0;
                o.win_hi = kBlockSize + kSentinel + kWinSlack;
            }
            o.items_cap = chunk_mode ? kChunkItems : (uint32_t)(len < kChunkItems ? (len ? len : 1) : kChunkItems);
            o.payload_cap = (uint32_t)align_up((size_t)o.items_cap * 5 + 16384, 16);
            // zero-initialised state: window, nodes, words (heads are 0xff, handled separately)
            o.zero_lo = plan.take(0, 256);
            o.win_alloc = (o.win_lo == 0 ? kSentinel : 0) + (size_t)(o.win_hi - o.win_lo);
            o.win = plan.take(o.win_alloc, 256);
            o.nodes = o.heads = o.klist = o.kcnt = o.lm = o.gext = 0;
            if (v7) {
                o.kcnt = plan.take(sizeof(uint32_t) * kNumBuckets * kHeadsStride);
                o.lm = plan.take((size_t)kNumBuckets * kNodesStride);
            } else {
                o.nodes = plan.take(sizeof(ENode) * kNumBuckets * kNodesStride);
            }
            o.words = plan.take(65536);
            o.zero_hi = plan.take(0, 256);
            if (v7) {
                o.klist = plan.take(sizeof(uint2) * kNumBuckets * kHeadsStride * 64);
                o.gext = plan.take(sizeof(uint32_t) * 512 * 48);
            } else {
                o.heads = plan.take(sizeof(int16_t) * kNumBuckets * kHeadsStride);
            }
            o.symranks = plan.take(sizeof(SymRank) * kNumCtx);
            o.items = plan.take(sizeof(Item) * o.items_cap);
            o.items_alt = v7 && !chunk_mode ? plan.take(sizeof(Item) * o.items_cap) : 0;
            o.sorted = plan.take(sizeof(uint32_t) * o.items_cap);
            o.payload = plan.take(o.payload_cap);
        } else {
            o.zero_lo = plan.take(0, 256);
            // stream mode decodes in place: 256 zero bytes precede out[0]; chunk mode keeps the lib.rs:102 window
            const size_t out_bytes = chunk_mode ? (size_t)kSentinel + kBlockSize + kSentinel + 64 : 256 + out_cap[i] + 64;
            o.out = plan.take(out_bytes, 256);
            o.nodes = plan.take(sizeof(Node) * kNumBuckets * kNodesStride);
            o.words = plan.take(65536);
            o.zero_hi = plan.take(0, 256);
            o.symranks = plan.take(sizeof(SymRank) * kNumCtx);
        }
    }
    b->arena_size = align_up(plan.off, 256);
    if (cudaMalloc(&b->arena, b->arena_size) != cudaSuccess) {
        cudaGetLastError();
        *err = ORZ_ENOMEM;
        delete b;
        return nullptr;
    }
    bool ok2 = true;
    if (mode == 0 && v7 && !chunk_mode) {
        ok2 = cudaStreamCreateWithFlags(&b->st2, cudaStreamNonBlocking) == cudaSuccess;
        for (int i = 0; i < 2 && ok2; i++)
            ok2 = cudaEventCreateWithFlags(&b->ev_parse[i], cudaEventDisableTiming) == cudaSuccess &&
                  cudaEventCreateWithFlags(&b->ev_emit[i], cudaEventDisableTiming) == cudaSuccess;
    }
    if (!ok2 || cudaStreamCreateWithFlags(&b->st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMallocHost(&b->h_status_pinned, sizeof(int32_t) * 2 * n) != cudaSuccess) {
        *err = ORZ_ECUDA;
        batch_destroy(b);
        return nullptr;
    }
    uint8_t *A = b->arena;
    if (mode == 0) {
        b->d_enc = reinterpret_cast<EncStream *>(A + ctl_off);
        b->enc.resize(n);
        for (int i = 0; i < n; i++) {
            const Offs &o = offs[i];
            EncStream &S = b->enc[i];
            memset(&S, 0, sizeof S);
            S.in = A + o.in;
            S.in_len = 0;
            S.out = A + o.out;
            S.out_cap = out_cap[i];
            S.win = (o.win_lo == 0) ? (A + o.win + kSentinel) : (A + o.win - o.win_lo);
            S.win_lo = o.win_lo;
            S.win_hi = o.win_hi;
            S.v7 = v7 ? 1u : 0u;
            S.nodes = v7 ? nullptr : reinterpret_cast<ENode *>(A + o.nodes);
            S.heads = v7 ? nullptr : reinterpret_cast<int16_t *>(A + o.heads);
            S.klist = v7 ? reinterpret_cast<uint2 *>(A + o.klist) : nullptr;
            S.kcnt = v7 ? reinterpret_cast<uint32_t *>(A + o.kcnt) : nullptr;
            S.lm = v7 ? A + o.lm : nullptr;
            S.gext = v7 ? reinterpret_cast<uint32_t *>(A + o.gext) : nullptr;
            S.words = reinterpret_cast<uint16_t *>(A + o.words);
            S.symranks = reinterpret_cast<SymRank *>(A + o.symranks);
            S.items = reinterpret_cast<Item *>(A + o.items);
            S.items_alt = v7 && !chunk_mode ? reinterpret_cast<Item *>(A + o.items_alt) : nullptr;
            S.sorted = reinterpret_cast<uint32_t *>(A + o.sorted);
            S.payload = A + o.payload;
            S.payload_cap = o.payload_cap;
            S.items_cap = o.items_cap;
            S.depth0 = (uint32_t)(cfg->match_depth > 0x7fffffffu ? 0x7fffffffu : cfg->match_depth);
            S.depth1 = (uint32_t)(cfg->lazy_match_depth1 > 0x7fffffffu ? 0x7fffffffu : cfg->lazy_match_depth1);
            S.depth2 = (uint32_t)(cfg->lazy_match_depth2 > 0x7fffffffu ? 0x7fffffffu : cfg->lazy_match_depth2);
            S.self_driven = chunk_mode ? 0 : 1;
            S.dbg_flags = getenv("ORZ_B200_DBG") ? (uint32_t)atoi(getenv("ORZ_B200_DBG")) : 0u;
            S.dbg_pos = getenv("ORZ_B200_DBG_POS") ? (uint32_t)strtoul(getenv("ORZ_B200_DBG_POS"), nullptr, 10) : 0u;
            S.dbg_inpos = getenv("ORZ_B200_DBG_INPOS") ? strtoull(getenv("ORZ_B200_DBG_INPOS"), nullptr, 10) : 0ull;
            S.dbgx_n = 0;
            S.prov = nullptr;
            S.prov_cap = S.prov_n = 0;
#ifndef ORZ_SIMT_EMU
            if (i == 0 && v7 && getenv("ORZ_B200_PROV")) { // debugging aid of the A/B builds: provenance of every parse step (stream 0 only)
                const size_t cap = in_len[i] + 1024;
                void *p = nullptr;
                if (cudaMalloc(&p, cap * 64) == cudaSuccess) {
                    cudaMemset(p, 0, cap * 64);
                    S.prov = reinterpret_cast<uint4 *>(p);
                    S.prov_cap = (uint32_t)cap;
                    b->prov = p;
                }
            }
#endif
            S.first_block = 1;   // lz.rs:63
            S.after_literal = 1; // lz.rs:64
            b->d_in[i] = A + o.in;
            b->d_out[i] = A + o.out;
            b->zero_lo[i] = A + o.zero_lo;
            b->zero_hi[i] = A + o.zero_hi;
            b->heads[i] = S.heads;
        }
    } else {
        b->d_dec = reinterpret_cast<DecStream *>(A + ctl_off);
        b->dec.resize(n);
        for (int i = 0; i < n; i++) {
            const Offs &o = offs[i];
            DecStream &S = b->dec[i];
            memset(&S, 0, sizeof S);
            S.in = A + o.in;
            S.in_len = 0;
            S.out = chunk_mode ? nullptr : A + o.out + 256;
            S.out_cap = out_cap[i];
            S.win = chunk_mode ? A + o.out + kSentinel : nullptr;
            S.nodes = reinterpret_cast<Node *>(A + o.nodes);
            S.words = reinterpret_cast<uint16_t *>(A + o.words);
            S.symranks = reinterpret_cast<SymRank *>(A + o.symranks);
            S.self_driven = chunk_mode ? 0 : 1;
            S.first_block = 1;
            S.after_literal = 1;
            S.spos = kPrematch;
            b->d_in[i] = A + o.in;
            b->d_out[i] = chunk_mode ? S.win : S.out;
            b->zero_lo[i] = A + o.zero_lo;
            b->zero_hi[i] = A + o.zero_hi;
            b->heads[i] = nullptr;
        }
    }
    return b;
}

// LZEncoder::new / LZDecoder::new + buffer allocation of lib.rs:67-70,102-105: everything zero, heads -1
static int batch_reset_state(orz_batch *b) {
    for (int i = 0; i < b->n; i++) {
        CU_TRY(cudaMemsetAsync(b->zero_lo[i], 0, (size_t)(b->zero_hi[i] - b->zero_lo[i]), b->st));
        if (b->mode == 0 && !b->v7)
            CU_TRY(cudaMemsetAsync(b->heads[i], 0xff, sizeof(int16_t) * kNumBuckets * kHeadsStride, b->st));
    }
    if (b->mode == 0) {
        for (int i = 0; i < b->n; i++) b->enc[i].in_len = b->in_len[i];
        CU_TRY(cudaMemcpyAsync(b->d_enc, b->enc.data(), sizeof(EncStream) * b->n, cudaMemcpyHostToDevice, b->st));
    } else {
        for (int i = 0; i < b->n; i++) b->dec[i].in_len = b->in_len[i];
        CU_TRY(cudaMemcpyAsync(b->d_dec, b->dec.data(), sizeof(DecStream) * b->n, cudaMemcpyHostToDevice, b->st));
    }
    return ORZ_OK;
}

static int fetch_status(orz_batch *b, bool *all_done) {
    const void *src = b->mode == 0 ? (const void *)&b->d_enc[0].status : (const void *)&b->d_dec[0].status;
    const size_t pitch = b->mode == 0 ? sizeof(EncStream) : sizeof(DecStream);
    // encode: {status, parse_done} are adjacent — a stream whose stage A reached the end of its input needs no further round
    const size_t width = b->mode == 0 ? 2 * sizeof(int32_t) : sizeof(int32_t);
    CU_TRY(cudaMemcpy2DAsync(b->h_status_pinned, width, src, pitch, width, b->n, cudaMemcpyDeviceToHost, b->st));
    CU_TRY(cudaStreamSynchronize(b->st));
    *all_done = true;
    for (int i = 0; i < b->n; i++) {
        b->status[i] = b->mode == 0 ? b->h_status_pinned[2 * i] : b->h_status_pinned[i];
        const bool parsed = b->mode == 0 && b->h_status_pinned[2 * i + 1] != 0;
        if (b->status[i] == 0 && !parsed) *all_done = false;
    }
    return ORZ_OK;
}

extern "C" orz_batch *orz_batch_create(int device, int mode, const orz_lzcfg *cfg, int nstreams, const size_t *in_len,
                                       const size_t *out_cap) {
    int err;
    if (!in_len || !out_cap || (mode != 0 && mode != 1)) return nullptr;
    return batch_create(device, mode, cfg, nstreams, in_len, out_cap, false, &err);
}
extern "C" void orz_batch_free(orz_batch *b) { batch_destroy(b); }

extern "C" int orz_batch_upload(orz_batch *b, int i, const uint8_t *host, size_t n) {
    if (!b || i < 0 || i >= b->n || n > b->in_cap[i] || (n && !host)) return ORZ_EARG;
    DEV_GUARD(b->device);
    if (n) CU_TRY(cudaMemcpyAsync(b->d_in[i], host, n, cudaMemcpyHostToDevice, b->st));
    CU_TRY(cudaMemsetAsync(b->d_in[i] + n, 0, 64, b->st));
    b->in_len[i] = n;
    return ORZ_OK;
}

extern "C" int orz_batch_run(orz_batch *b) {
    if (!b) return ORZ_EARG;
    DEV_GUARD(b->device);
    b->timer.reset();
    const int whole = b->timer.begin(4, b->st); // class 4: the whole run, reset included
    int rc = batch_reset_state(b);
    if (rc) return rc;
    bool all_done = false;
    if (b->mode == 0) {
        int id = b->timer.begin(2, b->st);
        launch_enc_begin(b->d_enc, b->n, b->st);
        b->timer.end(id, b->st);
        for (int k = 0;; k++) {
            if ((rc = fetch_status(b, &all_done))) return rc; // (waits for stage A of the previous round, not for its stages B-D)
            if (all_done) break;
            const int buf = b->v7 ? (k & 1) : 0;
            cudaStream_t se = b->v7 ? b->st2 : b->st;
            if (b->v7 && k >= 2) CU_TRY(cudaStreamWaitEvent(b->st, b->ev_emit[buf], 0)); // the item buffer is free again
            id = b->timer.begin(0, b->st);
            launch_parse(b->d_enc, b->n, b->v7, buf, b->st);
            b->timer.end(id, b->st);
            if (b->v7) { // the lib.rs:72-89 decision and, at the end of a block, slide + forward + next block's load (grid-wide)
                id = b->timer.begin(2, b->st);
                launch_chunk_done(b->d_enc, b->n, buf, b->st);
                launch_enc_advance(b->d_enc, b->n, b->st);
                b->timer.end(id, b->st);
                CU_TRY(cudaEventRecord(b->ev_parse[buf], b->st));
                CU_TRY(cudaStreamWaitEvent(se, b->ev_parse[buf], 0));
            }
            id = b->timer.begin(1, se);
            launch_emit(b->d_enc, b->n, buf, se);
            b->timer.end(id, se);
            if (b->v7) CU_TRY(cudaEventRecord(b->ev_emit[buf], se));
            CU_TRY(cudaGetLastError());
        }
        if (b->v7) { // the last chunks' stages B-D, then make `st` wait for them too (the timing events of the whole run live there)
            CU_TRY(cudaStreamSynchronize(b->st2));
        }
        // result lengths
        std::vector<EncStream> tmp(b->n);
        CU_TRY(cudaMemcpyAsync(tmp.data(), b->d_enc, sizeof(EncStream) * b->n, cudaMemcpyDeviceToHost, b->st));
        CU_TRY(cudaStreamSynchronize(b->st));
        for (int i = 0; i < b->n; i++) b->result_len[i] = tmp[i].out_pos;
        if (b->v7 && parse7_stats_enabled()) {
            unsigned long long d[8] = {0}, c[24] = {0};
            for (int i = 0; i < b->n; i++) {
                for (int k = 0; k < 8; k++) d[k] += tmp[i].dbg[k];
                for (int k = 0; k < 24; k++) c[k] += tmp[i].dbg7[k];
            }
            const double st = d[0] ? (double)d[0] : 1.0, wn = d[1] ? (double)d[1] : 1.0;
            fprintf(stderr,
                    "orz_b200 stage A v7: steps %llu windows %llu (%.2f steps/window) polls-not-ready %llu repairs %llu (%.4f/step: %llu records re-resolved — %llu alias, %llu one newer node, %llu in full — %llu rebuilt from HBM) "
                    "builds %llu (%.2f/step, %.0f cycles each), refreshes %llu (parse waited for one %llu times); list lag %.1f steps / %.0f bytes; "
                    "resolver cycles/step: total %.1f = team %.1f + chase %.1f + hand-over %.1f + repair %.1f + wait %.1f; verification cuts %llu; rounds ended by: block end / item cap %llu, repair %llu, record not built %llu, edge record %llu; "
                    "latency probes (cycles): 32 dependent IMADs %.0f, 16 dependent SHFLs %.0f, 16 dependent LDS %.0f; records seen with their stamp before their words: %llu by the decide team, %llu by repairs\n",
                    d[0], d[1], (double)d[0] / wn, d[2], d[3], d[3] / st, c[6], c[21], c[22], c[23], c[7], d[4], d[4] / st, d[4] ? (double)c[8] / d[4] : 0.0, d[5], d[6], c[9] / wn, c[10] / wn,
                    c[0] / st, c[1] / st, c[2] / st, c[3] / st, c[4] / st, c[5] / st, c[11], c[16] + c[19], c[17], c[18], c[20], c[15] ? (double)c[12] / c[15] : 0.0,
                    c[15] ? (double)c[13] / c[15] : 0.0, c[15] ? (double)c[14] / c[15] : 0.0, d[7] & 0xffffull, d[7] >> 16);
        }
        if (b->v7 && getenv("ORZ_B200_TORN")) { // records seen with their stamp before their words (stage A v7 waits for them; see sumcheck7)
            unsigned long long t = 0;
            for (int i = 0; i < b->n; i++) t += tmp[i].dbg[7];
            fprintf(stderr, "orz_b200 stage A v7: incomplete records seen: %llu by the decide team, %llu by repairs\n", t & 0xffffull, t >> 16);
        }
        if (b->v7 && getenv("ORZ_B200_DBG_POS")) {
            for (int i = 0; i < b->n; i++) { // the event trace of ORZ_B200_DBG_POS: tag | words << 24, then the words
                const uint32_t nx = tmp[i].dbgx_n < 1024u ? tmp[i].dbgx_n : 1024u;
                for (uint32_t k = 0; k < nx;) {
                    const uint32_t tag = tmp[i].dbgx[k] & 0xffffffu, nw = tmp[i].dbgx[k] >> 24;
                    fprintf(stderr, "orz_b200 trace: %s", tag == 1 ? "build" : tag == 2 ? "repair" : tag == 3 ? "hand-over" : tag == 4 ? "decide" : tag == 5 ? "candidates" : "?");
                    for (uint32_t w = 1; w <= nw && k + w < 1024u; w++) fprintf(stderr, " %u", tmp[i].dbgx[k + w]);
                    fprintf(stderr, "\n");
                    k += nw + 1;
                }
            }
        }
        if (!(b->v7 && parse7_stats_enabled()) && getenv("ORZ_B200_STATS")) {
            unsigned long long d[8] = {0};
            for (int i = 0; i < b->n; i++)
                for (int k = 0; k < 8; k++) d[k] += tmp[i].dbg[k];
            fprintf(stderr, "orz_b200 stage A: steps %llu, executor spins own %llu lazy1 %llu lazy2 %llu, builds %llu, aborted %llu, seqlock-failed %llu; own-probe not ready at first poll: building %llu, other-q %llu, stale-version %llu\n", d[0], d[1],
                    d[2], d[3], d[4], d[5] & 0xffffffffull, d[5] >> 32, d[6] >> 32, d[6] & 0xffffffffull, d[7]);
        }
    } else {
        int id = b->timer.begin(3, b->st);
        launch_decode(b->d_dec, b->n, b->st);
        b->timer.end(id, b->st);
        CU_TRY(cudaGetLastError());
        std::vector<DecStream> tmp(b->n);
        CU_TRY(cudaMemcpyAsync(tmp.data(), b->d_dec, sizeof(DecStream) * b->n, cudaMemcpyDeviceToHost, b->st));
        CU_TRY(cudaStreamSynchronize(b->st));
        for (int i = 0; i < b->n; i++) {
            b->result_len[i] = tmp[i].out_len;
            b->status[i] = tmp[i].status;
        }
        if (getenv("ORZ_B200_STATS")) {
            unsigned long long it = 0, ms = 0;
            for (int i = 0; i < b->n; i++) it += tmp[i].dbg[0], ms += tmp[i].dbg[1];
            fprintf(stderr, "orz_b200 decode: items %llu, rank-table cache misses %llu\n", it, ms);
        }
    }
    b->timer.end(whole, b->st);
    CU_TRY(cudaStreamSynchronize(b->st));
    b->timer.collect();
    b->ran = true;
    for (int i = 0; i < b->n; i++)
        if (b->status[i] < 0) return b->status[i];
    return ORZ_OK;
}

extern "C" int orz_batch_result_len(orz_batch *b, int i, size_t *len) {
    if (!b || i < 0 || i >= b->n || !len || !b->ran) return ORZ_EARG;
    if (b->status[i] < 0) return b->status[i];
    *len = (size_t)b->result_len[i];
    return ORZ_OK;
}

extern "C" int orz_batch_download(orz_batch *b, int i, uint8_t *host, size_t cap, size_t *len) {
    if (!b || i < 0 || i >= b->n || !b->ran) return ORZ_EARG;
    if (b->status[i] < 0) return b->status[i];
    const size_t n = (size_t)b->result_len[i];
    if (n > cap) return ORZ_ENOSPC;
    DEV_GUARD(b->device);
    if (n) CU_TRY(cudaMemcpyAsync(host, b->d_out[i], n, cudaMemcpyDeviceToHost, b->st));
    CU_TRY(cudaStreamSynchronize(b->st));
    if (len) *len = n;
    return ORZ_OK;
}

extern "C" void *orz_batch_result_dev(orz_batch *b, int i) {
    if (!b || i < 0 || i >= b->n) return nullptr;
    return b->d_out[i];
}

// debugging aid (not part of include/orz_b200.h): the four uint4 of provenance that stage A left for input position `index` of stream 0
extern "C" int orz_b200_debug_prov(orz_batch *b, size_t index, uint32_t *out16) {
    if (!b || !b->prov || !out16) return ORZ_EARG;
    DevGuard g(b->device);
    return cudaMemcpy(out16, reinterpret_cast<const uint8_t *>(b->prov) + index * 64, 64, cudaMemcpyDeviceToHost) == cudaSuccess ? ORZ_OK : ORZ_ECUDA;
}

extern "C" int orz_batch_timing(orz_batch *b, int cls, float *total_ms, int *launches) {
    if (!b || cls < 0 || cls > 4) return 0;
    if (total_ms) *total_ms = b->timer.ms[cls];
    if (launches) *launches = b->timer.launches[cls];
    return b->timer.launches[cls];
}

// ------------------------------------------------------------------------------------------------
// stream / batch entry points with host buffers (lib.rs:58-92, 94-129)
// ------------------------------------------------------------------------------------------------
// Host-buffer calls keep their last arena: allocating (and freeing) tens of MB of ring state per stream costs more
// than the kernels for short inputs.  One cached batch per process, reused when the shape of the call repeats.
#include <mutex>
namespace {
struct BatchCache {
    std::mutex mu;
    orz_batch *b = nullptr;
    int device = -1, mode = -1, n = 0;
    orz_lzcfg cfg{};
    std::vector<size_t> in_cap, out_cap;
    ~BatchCache() {
        if (b) batch_destroy(b);
    }
} g_caches[16]; // one cached batch (and one lock) per device: calls on different GPUs neither serialise nor evict each other
} // namespace

static int run_host_batch(int device, int mode, const orz_lzcfg *cfg, int n, const uint8_t *const *in, const size_t *in_len,
                          uint8_t *const *out, const size_t *out_cap, size_t *out_len) {
    if (n <= 0 || !in || !in_len || !out || !out_cap || !out_len) return ORZ_EARG;
    if (device < 0 || device >= 16) return ORZ_ECUDA;
    BatchCache &g_cache = g_caches[device];
    std::lock_guard<std::mutex> lock(g_cache.mu);
    bool reuse = g_cache.b && g_cache.device == device && g_cache.mode == mode && g_cache.n == n &&
                 (mode != 0 || (cfg->match_depth == g_cache.cfg.match_depth && cfg->lazy_match_depth1 == g_cache.cfg.lazy_match_depth1 &&
                                cfg->lazy_match_depth2 == g_cache.cfg.lazy_match_depth2));
    for (int i = 0; reuse && i < n; i++)
        // same window class (one-block or full) and enough room
        reuse = in_len[i] <= g_cache.in_cap[i] && out_cap[i] <= g_cache.out_cap[i] &&
                (mode != 0 || ((in_len[i] <= kBlockNew) == (g_cache.in_cap[i] <= kBlockNew) && in_len[i] * 2 >= g_cache.in_cap[i]));
    if (!reuse) {
        if (g_cache.b) batch_destroy(g_cache.b);
        g_cache.b = nullptr;
        int err;
        std::vector<size_t> dev_cap(out_cap, out_cap + n);
        orz_batch *nb = batch_create(device, mode, cfg, n, in_len, dev_cap.data(), false, &err);
        if (!nb) return err;
        g_cache.b = nb;
        g_cache.device = device;
        g_cache.mode = mode;
        g_cache.n = n;
        if (cfg) g_cache.cfg = *cfg;
        g_cache.in_cap.assign(in_len, in_len + n);
        g_cache.out_cap.assign(out_cap, out_cap + n);
    }
    orz_batch *b = g_cache.b;
    int rc = ORZ_OK;
    for (int i = 0; i < n && rc == ORZ_OK; i++) rc = orz_batch_upload(b, i, in[i], in_len[i]);
    if (rc == ORZ_OK) rc = orz_batch_run(b);
    if (rc == ORZ_OK)
        for (int i = 0; i < n && rc == ORZ_OK; i++) rc = orz_batch_download(b, i, out[i], out_cap[i], &out_len[i]);
    if (rc != ORZ_OK) { // do not keep an arena in an unknown state
        batch_destroy(b);
        g_cache.b = nullptr;
    }
    return rc;
}

extern "C" int orz_encode_batch(int device, const orz_lzcfg *cfg, int nstreams, const uint8_t *const *in, const size_t *n,
                                uint8_t *const *out, const size_t *out_cap, size_t *out_len) {
    if (!cfg) return ORZ_EARG;
    return run_host_batch(device, 0, cfg, nstreams, in, n, out, out_cap, out_len);
}
extern "C" int orz_decode_batch(int device, int nstreams, const uint8_t *const *in, const size_t *n, uint8_t *const *out,
                                const size_t *out_cap, size_t *out_len) {
    return run_host_batch(device, 1, nullptr, nstreams, in, n, out, out_cap, out_len);
}
// Several devices: stream i -> devices[i % ndev], one host thread per device slot (each drives its own arena through the per-device cache)
#include <thread>
static int run_multi(const int *devices, int ndev, int mode, const orz_lzcfg *cfg, int n, const uint8_t *const *in, const size_t *in_len,
                     uint8_t *const *out, const size_t *out_cap, size_t *out_len) {
    if (!devices || ndev <= 0 || n <= 0 || !in || !in_len || !out || !out_cap || !out_len) return ORZ_EARG;
    std::vector<int> rc(ndev, ORZ_OK);
    std::vector<std::thread> th;
    for (int d = 0; d < ndev; d++) {
        th.emplace_back([&, d]() {
            std::vector<const uint8_t *> i2;
            std::vector<size_t> n2, c2, l2;
            std::vector<uint8_t *> o2;
            std::vector<int> idx;
            for (int i = d; i < n; i += ndev) {
                idx.push_back(i);
                i2.push_back(in[i]);
                n2.push_back(in_len[i]);
                o2.push_back(out[i]);
                c2.push_back(out_cap[i]);
            }
            if (idx.empty()) return;
            l2.resize(idx.size());
            rc[d] = run_host_batch(devices[d], mode, cfg, (int)idx.size(), i2.data(), n2.data(), o2.data(), c2.data(), l2.data());
            if (rc[d] == ORZ_OK)
                for (size_t k = 0; k < idx.size(); k++) out_len[idx[k]] = l2[k];
        });
    }
    for (auto &t : th) t.join();
    for (int d = 0; d < ndev; d++)
        if (rc[d] != ORZ_OK) return rc[d];
    return ORZ_OK;
}
extern "C" int orz_encode_batch_multi(const int *devices, int ndev, const orz_lzcfg *cfg, int nstreams, const uint8_t *const *in, const size_t *n,
                                      uint8_t *const *out, const size_t *out_cap, size_t *out_len) {
    if (!cfg) return ORZ_EARG;
    return run_multi(devices, ndev, 0, cfg, nstreams, in, n, out, out_cap, out_len);
}
extern "C" int orz_decode_batch_multi(const int *devices, int ndev, int nstreams, const uint8_t *const *in, const size_t *n, uint8_t *const *out,
                                      const size_t *out_cap, size_t *out_len) {
    return run_multi(devices, ndev, 1, nullptr, nstreams, in, n, out, out_cap, out_len);
}
extern "C" int orz_encode_stream(int device, const orz_lzcfg *cfg, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap,
                                 size_t *out_len) {
    return orz_encode_batch(device, cfg, 1, &in, &n, &out, &out_cap, out_len);
}
extern "C" int orz_decode_stream(int device, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap, size_t *out_len) {
    return orz_decode_batch(device, 1, &in, &n, &out, &out_cap, out_len);
}

// ------------------------------------------------------------------------------------------------
// segment container (extension): independent orz streams + index + CRC32
// ------------------------------------------------------------------------------------------------
static uint32_t crc32_ieee(const uint8_t *p, size_t n) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    uint32_t c = 0xffffffffu;
    for (size_t i = 0; i < n; i++) c = table[(c ^ p[i]) & 0xffu] ^ (c >> 8);
    return c ^ 0xffffffffu;
}
static void put32(uint8_t *p, uint32_t v) { memcpy(p, &v, 4); }
static void put64(uint8_t *p, uint64_t v) { memcpy(p, &v, 8); }
static uint32_t get32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint64_t get64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static const size_t kSegHdr = 24, kSegEnt = 12;

extern "C" size_t orz_segmented_bound(size_t n, size_t segment_size) {
    if (segment_size == 0) return 0;
    const size_t nseg = (n + segment_size - 1) / segment_size;
    return kSegHdr + nseg * kSegEnt + nseg * orz_encode_bound(segment_size < n ? segment_size : n) + 16;
}

extern "C" int orz_encode_segmented(int device, const orz_lzcfg *cfg, int level_tag, const uint8_t *in, size_t n, size_t segment_size,
                                    uint8_t *out, size_t out_cap, size_t *out_len) {
    if (!cfg || (n && !in) || !out || !out_len || segment_size == 0 || segment_size > 0xffffffffull) return ORZ_EARG;
    const size_t nseg = (n + segment_size - 1) / segment_size;
    if (nseg > 0x7fffffffull) return ORZ_EARG;
    const size_t table_end = kSegHdr + nseg * kSegEnt;
    if (out_cap < table_end) return ORZ_ENOSPC;
    memcpy(out, "ORZS", 4);
    out[4] = 1;
    out[5] = (uint8_t)level_tag;
    out[6] = out[7] = 0;
    put64(out + 8, n);
    put32(out + 16, (uint32_t)segment_size);
    put32(out + 20, (uint32_t)nseg);
    if (nseg == 0) {
        *out_len = table_end;
        return ORZ_OK;
    }
    std::vector<const uint8_t *> ins(nseg);
    std::vector<size_t> lens(nseg), caps(nseg), got(nseg);
    std::vector<std::vector<uint8_t>> tmp(nseg);
    std::vector<uint8_t *> outs(nseg);
    for (size_t i = 0; i < nseg; i++) {
        ins[i] = in + i * segment_size;
        lens[i] = (i + 1 < nseg) ? segment_size : n - i * segment_size;
        caps[i] = orz_encode_bound(lens[i]);
        tmp[i].resize(caps[i]);
        outs[i] = tmp[i].data();
    }
    const int rc = orz_encode_batch(device, cfg, (int)nseg, ins.data(), lens.data(), outs.data(), caps.data(), got.data());
    if (rc != ORZ_OK) return rc;
    size_t pos = table_end;
    for (size_t i = 0; i < nseg; i++) {
        if (pos + got[i] > out_cap) return ORZ_ENOSPC;
        uint8_t *ent = out + kSegHdr + i * kSegEnt;
        put32(ent, (uint32_t)lens[i]);
        put32(ent + 4, (uint32_t)got[i]);
        put32(ent + 8, crc32_ieee(ins[i], lens[i]));
        memcpy(out + pos, outs[i], got[i]);
        pos += got[i];
    }
    *out_len = pos;
    return ORZ_OK;
}

extern "C" int orz_segmented_raw_len(const uint8_t *in, size_t n, size_t *raw_len) {
    if (!in || !raw_len || n < kSegHdr || memcmp(in, "ORZS", 4) != 0 || in[4] != 1) return ORZ_EINVAL;
    *raw_len = (size_t)get64(in + 8);
    return ORZ_OK;
}

extern "C" int orz_decode_segmented(int device, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap, size_t *out_len) {
    size_t raw = 0;
    if (!out_len || orz_segmented_raw_len(in, n, &raw) != ORZ_OK) return ORZ_EINVAL;
    const size_t seg_size = get32(in + 16), nseg = get32(in + 20);
    if (nseg > 0x7fffffffull || nseg > (n - kSegHdr) / kSegEnt) return ORZ_EINVAL; // untrusted count: bounded by the table it must fit in
    const size_t table_end = kSegHdr + nseg * kSegEnt;
    if (table_end > n || (nseg && seg_size == 0) || raw > nseg * (uint64_t)seg_size) return ORZ_EINVAL;
    if (raw > out_cap) return ORZ_ENOSPC;
    if (nseg == 0) {
        *out_len = 0;
        return raw == 0 ? ORZ_OK : ORZ_EINVAL;
    }
    std::vector<const uint8_t *> ins(nseg);
    std::vector<size_t> lens(nseg), caps(nseg), got(nseg);
    std::vector<uint8_t *> outs(nseg);
    size_t pos = table_end, opos = 0;
    for (size_t i = 0; i < nseg; i++) {
        const uint8_t *ent = in + kSegHdr + i * kSegEnt;
        const size_t rl = get32(ent), cl = get32(ent + 4);
        if (rl > seg_size || pos + cl > n || opos + rl > raw) return ORZ_EINVAL;
        ins[i] = in + pos;
        lens[i] = cl;
        outs[i] = out + opos;
        caps[i] = rl;
        pos += cl;
        opos += rl;
    }
    if (opos != raw) return ORZ_EINVAL;
    const int rc = orz_decode_batch(device, (int)nseg, ins.data(), lens.data(), outs.data(), caps.data(), got.data());
    if (rc != ORZ_OK) return rc == ORZ_ENOSPC ? ORZ_EINVAL : rc; // a segment larger than its index entry is a corrupt container
    for (size_t i = 0; i < nseg; i++) {
        const uint8_t *ent = in + kSegHdr + i * kSegEnt;
        if (got[i] != caps[i] || crc32_ieee(outs[i], got[i]) != get32(ent + 8)) return ORZ_EINVAL;
    }
    *out_len = raw;
    return ORZ_OK;
}

// ------------------------------------------------------------------------------------------------
// operator interface (lz.rs:75-95, 353-366): host drives chunks, codec state stays in HBM
// ------------------------------------------------------------------------------------------------
struct orz_lzenc {
    orz_batch *b;
    const uint8_t *last_sbuf;
    size_t last_len, last_spos;
    bool started;
};
struct orz_lzdec {
    orz_batch *b;
    size_t in_cap;
    bool started;
};

extern "C" orz_lzenc *orz_lzenc_new(int device) {
    int err;
    orz_lzcfg cfg = {45, 27, 18};
    size_t in_len = 0, out_cap = 0;
    orz_batch *b = batch_create(device, 0, &cfg, 1, &in_len, &out_cap, true, &err);
    if (!b) return nullptr;
    if (batch_reset_state(b) != ORZ_OK || cudaStreamSynchronize(b->st) != cudaSuccess) {
        batch_destroy(b);
        return nullptr;
    }
    orz_lzenc *e = new (std::nothrow) orz_lzenc{b, nullptr, 0, 0, false};
    if (!e) batch_destroy(b);
    return e;
}
extern "C" void orz_lzenc_free(orz_lzenc *e) {
    if (!e) return;
    batch_destroy(e->b);
    delete e;
}
extern "C" int orz_lzenc_forward(orz_lzenc *e, size_t forward_len) {
    if (!e) return ORZ_EARG;
    orz_batch *b = e->b;
    DEV_GUARD(b->device);
    launch_enc_forward(b->d_enc, (uint32_t)(forward_len > 0x1ffffffu ? 0x1ffffffu : forward_len), b->v7, b->st);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(b->st));
    e->last_sbuf = nullptr; // the caller slid its window (lib.rs:83): force a re-upload
    return ORZ_OK;
}
extern "C" int orz_lzenc_encode(orz_lzenc *e, const orz_lzcfg *cfg, const uint8_t *sbuf, size_t sbuf_len, uint8_t *tbuf,
                                size_t tbuf_cap, size_t spos, size_t *spos_out, size_t *tlen_out) {
    if (!e || !cfg || !sbuf || !tbuf || !spos_out || !tlen_out) return ORZ_EARG;
    if (sbuf_len > kBlockSize || spos >= sbuf_len || spos < 4) return ORZ_EARG;
    orz_batch *b = e->b;
    DEV_GUARD(b->device);
    EncStream &H = b->enc[0];
    const bool cont = e->started && sbuf == e->last_sbuf && sbuf_len == e->last_len && spos == e->last_spos && spos != kPrematch;
    if (!cont) { // (re)upload the window including both sentinels: those bytes are bit-visible
        CU_TRY(cudaMemcpyAsync(H.win - kSentinel, sbuf - kSentinel, sbuf_len + 2 * (size_t)kSentinel, cudaMemcpyHostToDevice, b->st));
    }
    // patch the control block: cfg, spos, sbuf_len
    struct Patch {
        uint32_t depth0, depth1, depth2;
    } p;
    p.depth0 = (uint32_t)(cfg->match_depth > 0x7fffffffu ? 0x7fffffffu : cfg->match_depth);
    p.depth1 = (uint32_t)(cfg->lazy_match_depth1 > 0x7fffffffu ? 0x7fffffffu : cfg->lazy_match_depth1);
    p.depth2 = (uint32_t)(cfg->lazy_match_depth2 > 0x7fffffffu ? 0x7fffffffu : cfg->lazy_match_depth2);
    CU_TRY(cudaMemcpyAsync(&b->d_enc[0].depth0, &p, sizeof p, cudaMemcpyHostToDevice, b->st));
    uint32_t ss[2] = {(uint32_t)spos, (uint32_t)sbuf_len};
    CU_TRY(cudaMemcpyAsync(&b->d_enc[0].spos, ss, sizeof ss, cudaMemcpyHostToDevice, b->st));
    if (b->v7 && (cfg->match_depth > kV7MaxDepth || cfg->lazy_match_depth1 > kV7MaxDepth || cfg->lazy_match_depth2 > kV7MaxDepth))
        return ORZ_EARG; // the key-list layout holds 48 candidates per position (main.rs:97-102 uses at most 45); ORZ_B200_LEGACY_STAGE_A=1 lifts this
    launch_parse(b->d_enc, 1, b->v7, 0, b->st);
    if (b->v7) launch_chunk_done(b->d_enc, 1, 0, b->st);
    launch_emit(b->d_enc, 1, 0, b->st);
    CU_TRY(cudaGetLastError());
    EncStream R;
    CU_TRY(cudaMemcpyAsync(&R, b->d_enc, sizeof R, cudaMemcpyDeviceToHost, b->st));
    CU_TRY(cudaStreamSynchronize(b->st));
    if (R.status < 0) return R.status;
    if (R.payload_len > tbuf_cap) return ORZ_ENOSPC;
    CU_TRY(cudaMemcpyAsync(tbuf, H.payload, R.payload_len, cudaMemcpyDeviceToHost, b->st));
    CU_TRY(cudaStreamSynchronize(b->st));
    *spos_out = R.spos;
    *tlen_out = R.payload_len;
    e->started = true;
    e->last_sbuf = sbuf;
    e->last_len = sbuf_len;
    e->last_spos = R.spos;
    return ORZ_OK;
}

extern "C" orz_lzdec *orz_lzdec_new(int device) {
    int err;
    size_t in_len = (size_t)kTbufLen, out_cap = kBlockSize; // lib.rs:103: payloads are < 3*PREMATCH bytes
    orz_batch *b = batch_create(device, 1, nullptr, 1, &in_len, &out_cap, true, &err);
    if (!b) return nullptr;
    if (batch_reset_state(b) != ORZ_OK || cudaStreamSynchronize(b->st) != cudaSuccess) {
        batch_destroy(b);
        return nullptr;
    }
    orz_lzdec *d = new (std::nothrow) orz_lzdec{b, in_len, false};
    if (!d) batch_destroy(b);
    return d;
}
extern "C" void orz_lzdec_free(orz_lzdec *d) {
    if (!d) return;
    batch_destroy(d->b);
    delete d;
}
extern "C" int orz_lzdec_forward(orz_lzdec *d, size_t forward_len) {
    if (!d) return ORZ_EARG;
    orz_batch *b = d->b;
    DEV_GUARD(b->device);
    if (forward_len == 0) return ORZ_OK;
    if (forward_len > kBlockSize) forward_len = kBlockSize;
    // mirror the caller's copy_within (lib.rs:121) on the device window, then rebase the ring (lz.rs:359-364)
    uint8_t *win = b->dec[0].win;
    const size_t keep = kBlockSize - forward_len;
    if (keep > 0) {
        if (keep <= forward_len) {
            CU_TRY(cudaMemcpyAsync(win, win + forward_len, keep, cudaMemcpyDeviceToDevice, b->st));
        } else { // overlapping ranges: stage through a temporary
            uint8_t *tmp = nullptr;
            CU_TRY(cudaMalloc(&tmp, keep));
            cudaError_t ce = cudaMemcpyAsync(tmp, win + forward_len, keep, cudaMemcpyDeviceToDevice, b->st);
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(win, tmp, keep, cudaMemcpyDeviceToDevice, b->st);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(b->st);
            cudaFree(tmp);
            CU_TRY(ce);
        }
    }
    launch_dec_forward(b->d_dec, (uint32_t)forward_len, b->st);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(b->st));
    return ORZ_OK;
}
extern "C" int orz_lzdec_decode(orz_lzdec *d, const uint8_t *tbuf, size_t tlen, uint8_t *sbuf, size_t sbuf_cap, size_t spos,
                                size_t *spos_end) {
    if (!d || !tbuf || !sbuf || !spos_end) return ORZ_EARG;
    if (tlen == 0 || tlen >= kTbufLen) return ORZ_EINVAL; // lib.rs:111-113
    if (spos < 4 || spos > kBlockSize || sbuf_cap < kBlockSize) return ORZ_EARG;
    orz_batch *b = d->b;
    DEV_GUARD(b->device);
    CU_TRY(cudaMemcpyAsync(b->d_in[0], tbuf, tlen, cudaMemcpyHostToDevice, b->st));
    CU_TRY(cudaMemsetAsync(b->d_in[0] + tlen, 0, 16, b->st));
    uint64_t il = tlen;
    CU_TRY(cudaMemcpyAsync(&b->d_dec[0].in_len, &il, sizeof il, cudaMemcpyHostToDevice, b->st));
    uint32_t sp = (uint32_t)spos;
    CU_TRY(cudaMemcpyAsync(&b->d_dec[0].spos, &sp, sizeof sp, cudaMemcpyHostToDevice, b->st));
    launch_decode(b->d_dec, 1, b->st);
    CU_TRY(cudaGetLastError());
    DecStream R;
    CU_TRY(cudaMemcpyAsync(&R, b->d_dec, sizeof R, cudaMemcpyDeviceToHost, b->st));
    CU_TRY(cudaStreamSynchronize(b->st));
    if (R.status < 0) {
        int32_t zero = 0; // leave the decoder usable for diagnostics; the stream itself is broken
        cudaMemcpyAsync(&b->d_dec[0].status, &zero, sizeof zero, cudaMemcpyHostToDevice, b->st);
        cudaStreamSynchronize(b->st);
        return R.status;
    }
    if (R.spos < spos || R.spos > kBlockSize) return ORZ_EINVAL;
    CU_TRY(cudaMemcpyAsync(sbuf + spos, b->dec[0].win + spos, R.spos - spos, cudaMemcpyDeviceToHost, b->st));
    CU_TRY(cudaStreamSynchronize(b->st));
    *spos_end = R.spos;
    return ORZ_OK;
}
