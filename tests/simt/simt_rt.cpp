// simt_rt.cpp — fibre scheduler of the SIMT emulator (see simt_rt.h; TEST INFRASTRUCTURE ONLY).
//
// Scheduling modes (environment):
//   default            FIFO: a fibre runs until it blocks in a collective / barrier or reaches a poll point
//   SIMT_SEED=n        seeded random choice of the next fibre at every poll point and wake-up (race shaking)
//   SIMT_REVERSE=1     FIFO, but the lanes of a warp leave every collective in DESCENDING lane order: code between two
//                      collectives that only works because lane 0 (or "the earlier lane") happens to run first fails
//   SIMT_VT=1          virtual-time order: always run the runnable fibre with the smallest virtual clock; the
//                      kernels advance the clock with ORZ_COST(cycles) at their dependent loads, so a spin loop
//                      observes a flag exactly when its writer's clock says it was written.  This turns the
//                      emulator into a discrete-event latency model of the block-internal protocols.
#include "simt_rt.h"

#include <sys/mman.h>
#include <time.h>

#include <deque>
#include <queue>
#include <string>
#include <vector>

extern "C" void simt_switch(void **from_sp, void *to_sp);
asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size simt_switch,.-simt_switch
)");

namespace simt {

Ctx *cur = nullptr;
int cur_site = 0;

namespace {

constexpr size_t kStackBytes = 256 * 1024;

enum State { FREE, RUNNABLE, WAIT_WARP, WAIT_BLOCK, WAIT_CLUSTER, WAIT_NAMED, DONE };

struct WarpState;
struct Block;

struct Fiber {
    void *sp = nullptr;
    uint8_t *stack = nullptr;
    Ctx ctx;
    int state = FREE;
    uint32_t res = 0;
    unsigned long long seq = 0;
    WarpState *w = nullptr;
    Block *blk = nullptr;
};

struct WarpState {
    uint32_t arrived = 0, mask = 0, exist = 0;
    int op = 0, site = 0;
    uint32_t val[32], param[32];
    Fiber *lanes[32];
};

struct NamedBar {
    int arrived = 0;
    std::vector<Fiber *> waiters;
};

struct Block {
    std::vector<Fiber *> fibers;
    std::vector<WarpState> warps;
    int live = 0;
    int bar_arrived = 0;
    NamedBar named[16];
    std::vector<uint8_t> smem;
};

struct Cluster {
    std::vector<Block> blocks;
    std::vector<uint8_t *> smem_ptrs;
    int live = 0;
    int bar_arrived = 0;
};

std::vector<Fiber *> g_pool;
Fiber g_main;
Fiber *g_cur = nullptr;
Cluster *g_cluster = nullptr;
const std::function<void()> *g_body = nullptr;
std::string g_error;
bool g_error_pending = false;

// scheduler state
bool g_vt_mode = false;
bool g_random = false;
bool g_reverse = false; // SIMT_REVERSE=1: lanes resume from a collective in descending lane order (exposes code that leans on lane 0 running first)
uint64_t g_rng = 0x9e3779b97f4a7c15ull;
unsigned long long g_seq = 0;
unsigned long long g_switches = 0, g_switch_limit = 0;
int g_poll_cost = 32, g_coll_cost = 24, g_bar_cost = 40;
std::deque<Fiber *> g_fifo;
struct VtCmp {
    bool operator()(const Fiber *a, const Fiber *b) const {
        if (a->ctx.vt != b->ctx.vt) return a->ctx.vt > b->ctx.vt;
        return a->seq > b->seq;
    }
};
std::priority_queue<Fiber *, std::vector<Fiber *>, VtCmp> g_heap;
long long g_vt_end = 0; // largest clock seen at thread exit (kernel duration in virtual cycles)

inline uint64_t rnd() {
    g_rng ^= g_rng << 13;
    g_rng ^= g_rng >> 7;
    g_rng ^= g_rng << 17;
    return g_rng;
}

inline bool queue_empty() { return g_vt_mode ? g_heap.empty() : g_fifo.empty(); }

inline void push_back(Fiber *f) {
    f->state = RUNNABLE;
    f->seq = ++g_seq;
    if (g_vt_mode) g_heap.push(f);
    else if (g_random && !g_fifo.empty()) g_fifo.insert(g_fifo.begin() + (long)(rnd() % (g_fifo.size() + 1)), f);
    else g_fifo.push_back(f);
}
inline void push_front(Fiber *f) {
    f->state = RUNNABLE;
    f->seq = ++g_seq;
    if (g_vt_mode) g_heap.push(f);
    else if (g_random && !g_fifo.empty()) g_fifo.insert(g_fifo.begin() + (long)(rnd() % (g_fifo.size() + 1)), f);
    else g_fifo.push_front(f);
}
inline Fiber *pop() {
    if (g_vt_mode) {
        if (g_heap.empty()) return nullptr;
        Fiber *f = g_heap.top();
        g_heap.pop();
        return f;
    }
    if (g_fifo.empty()) return nullptr;
    Fiber *f = g_fifo.front();
    g_fifo.pop_front();
    return f;
}

inline void switch_to(Fiber *next) {
    Fiber *prev = g_cur;
    if (next == prev) return; // (random / virtual-time order can pick the fibre that just yielded)
    g_cur = next;
    cur = &next->ctx;
    g_switches++;
    simt_switch(&prev->sp, next->sp);
}

void fail(const std::string &why) {
    if (g_error.empty()) g_error = why;
    g_error_pending = true;
}

// the current fibre cannot continue: run something else (or give control back to launch())
void block_current() {
    if (g_switch_limit && g_switches > g_switch_limit) {
        fail("SIMT emulator: switch limit exceeded (livelock?)");
        switch_to(&g_main);
    }
    Fiber *next = pop();
    if (!next) next = &g_main; // nothing runnable: finished or deadlocked, launch() decides
    switch_to(next);
}

long long release_block_barrier(Block *b) {
    long long mx = g_cur ? g_cur->ctx.vt : 0;
    for (Fiber *f : b->fibers)
        if (f->state == WAIT_BLOCK) mx = std::max(mx, f->ctx.vt);
    for (Fiber *f : b->fibers)
        if (f->state == WAIT_BLOCK) {
            f->ctx.vt = mx + g_bar_cost;
            push_back(f);
        }
    b->bar_arrived = 0;
    return mx;
}
void release_cluster_barrier(Cluster *c) {
    long long mx = 0;
    for (Block &b : c->blocks)
        for (Fiber *f : b.fibers)
            if (f->state == WAIT_CLUSTER || f == g_cur) mx = std::max(mx, f->ctx.vt);
    for (Block &b : c->blocks)
        for (Fiber *f : b.fibers)
            if (f->state == WAIT_CLUSTER) {
                f->ctx.vt = mx + 4 * g_bar_cost;
                push_back(f);
            }
    c->bar_arrived = 0;
}

void fiber_entry() {
    Fiber *f = g_cur;
    (*g_body)();
    f = g_cur;
    f->state = DONE;
    g_vt_end = std::max(g_vt_end, f->ctx.vt);
    Block *b = f->blk;
    b->live--;
    g_cluster->live--;
    f->w->exist &= ~(1u << f->ctx.lane);
    if (b->live > 0 && b->bar_arrived == b->live) release_block_barrier(b);
    if (g_cluster->live > 0 && g_cluster->bar_arrived == g_cluster->live) release_cluster_barrier(g_cluster);
    block_current();
    abort(); // a finished fibre is never resumed
}

Fiber *get_fiber(size_t i) {
    while (g_pool.size() <= i) {
        Fiber *f = new Fiber();
        f->stack = (uint8_t *)mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (f->stack == (uint8_t *)MAP_FAILED) {
            perror("simt: mmap");
            abort();
        }
        g_pool.push_back(f);
    }
    return g_pool[i];
}

void prime(Fiber *f) {
    uintptr_t top = ((uintptr_t)f->stack + kStackBytes) & ~(uintptr_t)15;
    void **s = (void **)top;
    s[-1] = nullptr;               // fake return address of fiber_entry (keeps the ABI stack alignment)
    s[-2] = (void *)&fiber_entry;  // popped by `ret` in simt_switch
    for (int i = 3; i <= 8; i++) s[-i] = nullptr; // rbp rbx r12 r13 r14 r15
    f->sp = (void *)(s - 8);
}

void read_env() {
    static bool done = false;
    if (done) return;
    done = true;
    if (const char *e = getenv("SIMT_VT")) g_vt_mode = atoi(e) != 0;
    if (const char *e = getenv("SIMT_SEED")) {
        g_random = true;
        g_rng ^= (uint64_t)strtoull(e, nullptr, 10) * 0x2545f4914f6cdd1dull + 1;
    }
    if (const char *e = getenv("SIMT_REVERSE")) g_reverse = atoi(e) != 0;
    if (const char *e = getenv("SIMT_SWITCH_LIMIT")) g_switch_limit = strtoull(e, nullptr, 10);
    if (const char *e = getenv("SIMT_POLL_COST")) g_poll_cost = atoi(e);
    if (const char *e = getenv("SIMT_COLL_COST")) g_coll_cost = atoi(e);
}

} // namespace

long long now() { return g_cur ? g_cur->ctx.vt : 0; }
void cost(int cycles) {
    if (g_cur) g_cur->ctx.vt += cycles;
}
const char *last_error() { return g_error.empty() ? nullptr : g_error.c_str(); }

void poll() {
    Fiber *f = g_cur;
    f->ctx.vt += g_poll_cost;
    if (queue_empty()) return;
    push_back(f);
    block_current();
}

uint32_t collective(int op, uint32_t mask, uint32_t v, uint32_t param) {
    Fiber *f = g_cur;
    WarpState *w = f->w;
    const unsigned lane = f->ctx.lane;
    const int site = cur_site;
    cur_site = 0;
    if (w->arrived == 0) {
        w->mask = mask;
        w->op = op;
        w->site = site;
    } else if (w->site != site) {
        fail("SIMT emulator: lanes of one warp reached a collective at different source sites (divergent branch around a "
             "__shfl/__ballot/__syncwarp: undefined behaviour on the GPU): source lines " + std::to_string(w->site / 4096) + " and " +
             std::to_string(site / 4096) + " (0 = a __reduce_*_sync, which carries no site)");
        switch_to(&g_main);
    } else if (w->mask != mask || w->op != op) {
        fail("SIMT emulator: lanes of one warp are in different collectives (divergent __shfl/__ballot masks)");
        switch_to(&g_main);
    }
    w->val[lane] = v;
    w->param[lane] = param;
    w->arrived |= 1u << lane;
    if (mask & ~(w->arrived | w->exist)) {
        fail("SIMT emulator: collective names a lane that has exited");
        switch_to(&g_main);
    }
    if ((w->arrived & mask) != mask) {
        f->state = WAIT_WARP;
        block_current();
        return f->res;
    }
    // last arrival: compute every participant's result
    uint32_t ball = 0, red = 0;
    long long mx = 0;
    bool first = true;
    for (unsigned l = 0; l < 32; l++) {
        if (!((mask >> l) & 1u)) continue;
        mx = std::max(mx, w->lanes[l]->ctx.vt);
        const uint32_t x = w->val[l];
        switch (op) {
        case OP_BALLOT: ball |= (x ? 1u : 0u) << l; break;
        case OP_MAX: red = first ? x : std::max(red, x); break;
        case OP_MIN: red = first ? x : std::min(red, x); break;
        case OP_ADD: red = first ? x : red + x; break;
        case OP_OR: red = first ? x : (red | x); break;
        case OP_XOR: red = first ? x : (red ^ x); break;
        case OP_AND: red = first ? x : (red & x); break;
        default: break;
        }
        first = false;
    }
    for (unsigned l = 0; l < 32; l++) {
        if (!((mask >> l) & 1u)) continue;
        Fiber *t = w->lanes[l];
        uint32_t r = 0;
        switch (op) {
        case OP_IDX: {
            const unsigned s = w->param[l] & 31u;
            r = ((mask >> s) & 1u) ? w->val[s] : w->val[l];
            break;
        }
        case OP_BALLOT: r = ball; break;
        case OP_MATCH_ANY:
            for (unsigned k = 0; k < 32; k++)
                if (((mask >> k) & 1u) && w->val[k] == w->val[l]) r |= 1u << k;
            break;
        case OP_SYNC: r = 0; break;
        default: r = red; break;
        }
        t->res = r;
        t->ctx.vt = mx + g_coll_cost;
    }
    w->arrived = 0;
    // the other lanes go to the BACK of the queue: a warp that spins on collectives only must not starve the rest
    for (unsigned k = 0; k < 32; k++) {
        const unsigned l = g_reverse ? 31 - k : k;
        if (!((mask >> l) & 1u) || l == lane) continue;
        push_back(w->lanes[l]);
    }
    if (g_vt_mode || g_random || g_reverse) { // let the scheduler order this lane with the rest
        push_back(f);
        block_current();
    }
    return f->res;
}

// bar.sync / bar.arrive on a named barrier (PTX barriers 1..15): `count` threads complete it; bar.sync blocks, bar.arrive does not
void named_barrier(int id, int count, bool sync) {
    Fiber *f = g_cur;
    NamedBar &nb = f->blk->named[id & 15];
    nb.arrived++;
    if (nb.arrived >= count) {
        long long mx = f->ctx.vt;
        for (Fiber *w : nb.waiters) mx = std::max(mx, w->ctx.vt);
        for (Fiber *w : nb.waiters) {
            w->ctx.vt = mx + g_bar_cost;
            push_back(w);
        }
        nb.waiters.clear();
        nb.arrived = 0;
        return;
    }
    if (sync) {
        f->state = WAIT_NAMED;
        nb.waiters.push_back(f);
        block_current();
    }
}

void syncthreads() {
    Fiber *f = g_cur;
    Block *b = f->blk;
    b->bar_arrived++;
    if (b->bar_arrived == b->live) {
        f->ctx.vt = release_block_barrier(b) + g_bar_cost;
        return;
    }
    f->state = WAIT_BLOCK;
    block_current();
}

void cluster_sync() {
    Fiber *f = g_cur;
    Cluster *c = g_cluster;
    c->bar_arrived++;
    if (c->bar_arrived == c->live) {
        release_cluster_barrier(c);
        f->ctx.vt += 4 * g_bar_cost;
        return;
    }
    f->state = WAIT_CLUSTER;
    block_current();
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body, unsigned cluster) {
    read_env();
    if (g_error_pending) return; // a previous launch of this sequence failed: the stream is dead, like CUDA
    const unsigned nthreads = block.x * block.y * block.z;
    const unsigned nblocks = grid.x * grid.y * grid.z;
    g_switches = 0;
    g_vt_end = 0;
    if (cluster == 0) cluster = 1;
    if (nthreads == 0 || nblocks == 0) return;
    // the launch limits of the target (sm_100a): what the GPU would refuse, the emulator refuses too
    if (nthreads > 1024 || smem > 227 * 1024 || cluster > 8 || nblocks % cluster != 0) {
        fail("SIMT emulator: launch exceeds the device limits (1024 threads, 227 KiB shared memory, clusters of 8, grid % cluster)");
        return;
    }
    g_body = &body;
    for (unsigned b0 = 0; b0 < nblocks; b0 += cluster) {
        Cluster C;
        const unsigned nb = std::min(cluster, nblocks - b0);
        C.blocks.resize(nb);
        C.smem_ptrs.resize(nb);
        g_cluster = &C;
        size_t fi = 0;
        while (!g_fifo.empty()) g_fifo.pop_front();
        while (!g_heap.empty()) g_heap.pop();
        for (unsigned r = 0; r < nb; r++) {
            Block &B = C.blocks[r];
            B.smem.assign(smem + 64, 0xcd); // shared memory starts undefined
            C.smem_ptrs[r] = (uint8_t *)(((uintptr_t)B.smem.data() + 15) & ~(uintptr_t)15);
        }
        for (unsigned r = 0; r < nb; r++) {
            Block &B = C.blocks[r];
            const unsigned bid = b0 + r;
            B.warps.resize((nthreads + 31) / 32);
            B.live = (int)nthreads;
            C.live += (int)nthreads;
            for (unsigned t = 0; t < nthreads; t++) {
                Fiber *f = get_fiber(fi++);
                f->blk = &B;
                f->w = &B.warps[t / 32];
                f->w->lanes[t % 32] = f;
                f->w->exist |= 1u << (t % 32);
                f->res = 0;
                Ctx &c = f->ctx;
                c.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                c.bid = dim3(bid % grid.x, (bid / grid.x) % grid.y, bid / (grid.x * grid.y));
                c.bdim = block;
                c.gdim = grid;
                c.lane = t % 32;
                c.warp = t / 32;
                c.cta_rank = r;
                c.cluster_size = nb;
                c.dyn_smem = C.smem_ptrs[r];
                c.cluster_smem = C.smem_ptrs.data();
                c.vt = 0;
                prime(f);
                B.fibers.push_back(f);
                f->state = RUNNABLE;
                f->seq = ++g_seq;
                if (g_vt_mode) g_heap.push(f);
                else g_fifo.push_back(f);
            }
        }
        g_cur = &g_main;
        block_current(); // returns here when nothing is runnable any more
        g_cur = nullptr;
        cur = nullptr;
        if (!g_error_pending && C.live != 0) {
            char msg[512];
            int blocked_w = 0, blocked_b = 0, blocked_c = 0;
            int ex_t = -1, ex_b = -1;
            for (Block &B : C.blocks)
                for (Fiber *f : B.fibers) {
                    if (f->state == WAIT_WARP) blocked_w++, ex_t = (int)(f->ctx.warp * 32 + f->ctx.lane), ex_b = (int)f->ctx.bid.x;
                    if (f->state == WAIT_BLOCK) blocked_b++;
                    if (f->state == WAIT_CLUSTER) blocked_c++;
                }
            snprintf(msg, sizeof msg,
                     "SIMT emulator: deadlock in block %u..%u: %d threads wait in a warp collective (e.g. block %d thread %d), %d at "
                     "__syncthreads, %d at a cluster barrier, none runnable",
                     b0, b0 + nb - 1, blocked_w, ex_b, ex_t, blocked_b, blocked_c);
            fail(msg);
        }
        g_cluster = nullptr;
        if (g_error_pending) {
            if (getenv("SIMT_VERBOSE")) fprintf(stderr, "%s\n", g_error.c_str());
            break;
        }
    }
    if (getenv("SIMT_VT_REPORT")) fprintf(stderr, "simt: launch of %u x %u threads ended at virtual cycle %lld\n", nblocks, nthreads, g_vt_end);
    g_body = nullptr;
}

long long vt_end() { return g_vt_end; }
void vt_reset() { g_vt_end = 0; }

} // namespace simt

cudaError_t cudaGetLastError() {
    if (simt::g_error_pending) {
        simt::g_error_pending = false;
        return cudaErrorLaunchFailure;
    }
    return cudaSuccess;
}

static double wall_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
cudaError_t cudaEventCreate(cudaEvent_t *e) {
    *e = new simt_event{0};
    return cudaSuccess;
}
cudaError_t cudaEventDestroy(cudaEvent_t e) {
    delete e;
    return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
    e->t = wall_ms();
    return cudaSuccess;
}
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = (float)(b->t - a->t);
    return cudaSuccess;
}
