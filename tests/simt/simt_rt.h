// simt_rt.h — host-side SIMT emulator used ONLY by the test suite (tests/simt/Makefile builds
// tests/simt/liborz_b200_emu.so from the very same .cu sources with -DORZ_SIMT_EMU).
//
// TEST INFRASTRUCTURE, like oracle/: the product library (orz_b200/liborz_b200.so) is built by nvcc for
// sm_100a and never includes this file.  The emulator lets the CPU-only test tier execute the real kernel
// source — builders, executor, seqlocks, shuffles, ballots, block barriers — and compare it byte for byte with
// the oracle, and it shakes out protocol races with a seeded random scheduler.  It says nothing about speed.
//
// Model: every CUDA thread is a fibre (own stack, cooperative switch, one OS thread).  Fibres switch only at
// warp collectives (__shfl*/__ballot/__syncwarp: block until every lane named in the mask has arrived),
// block barriers, and the volatile shared-memory helpers of the kernels (poll points).  Blocks of one launch
// run one after another, except that the CTAs of one thread-block cluster run together.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

// ---- qualifiers ---------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __restrict__ __restrict

// ---- vector types -------------------------------------------------------------------------------
struct __attribute__((aligned(16))) uint4 {
    unsigned x, y, z, w;
};
struct __attribute__((aligned(8))) uint2 {
    unsigned x, y;
};
struct __attribute__((aligned(16))) int4 {
    int x, y, z, w;
};
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace simt {
struct Ctx {
    dim3 tid, bid, bdim, gdim;
    unsigned lane, warp, cta_rank, cluster_size;
    uint8_t *dyn_smem;              // this CTA's dynamic shared memory
    uint8_t **cluster_smem;         // dyn_smem of every CTA of the cluster
    long long vt;                   // virtual time (cycles) of this thread, advanced by cost()
};
extern Ctx *cur;

enum Op { OP_SYNC, OP_IDX, OP_BALLOT, OP_MATCH_ANY, OP_MAX, OP_MIN, OP_ADD, OP_OR, OP_AND, OP_XOR };
uint32_t collective(int op, uint32_t mask, uint32_t v, uint32_t param);
extern int cur_site;                // source site of the collective being executed (see the macros at the end of this file)
void syncthreads();
void named_barrier(int id, int count, bool sync);
void cluster_sync();
void poll();                        // a poll point: lets other fibres run
void cost(int cycles);              // virtual-time bookkeeping (timing model)
long long now();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body, unsigned cluster = 1);
const char *last_error();           // non-null after a deadlock / watchdog abort
long long vt_end();                 // largest virtual clock at thread exit since vt_reset()
void vt_reset();
} // namespace simt

#define threadIdx (simt::cur->tid)
#define blockIdx (simt::cur->bid)
#define blockDim (simt::cur->bdim)
#define gridDim (simt::cur->gdim)
#define ORZ_DYN_SHARED(name) uint8_t *name = simt::cur->dyn_smem

// ---- warp intrinsics ----------------------------------------------------------------------------
template <class T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    const int lane = (int)simt::cur->lane;
    const int s = (lane & ~(width - 1)) | (src & (width - 1));
    if (sizeof(T) <= 4) {
        uint32_t u = 0;
        memcpy(&u, &v, sizeof(T));
        u = simt::collective(simt::OP_IDX, mask, u, (uint32_t)s);
        T r;
        memcpy(&r, &u, sizeof(T));
        return r;
    } else {
        static_assert(sizeof(T) <= 8, "shuffle of > 8 bytes");
        uint32_t u[2] = {0, 0};
        memcpy(u, &v, sizeof(T));
        u[0] = simt::collective(simt::OP_IDX, mask, u[0], (uint32_t)s);
        u[1] = simt::collective(simt::OP_IDX, mask, u[1], (uint32_t)s);
        T r;
        memcpy(&r, u, sizeof(T));
        return r;
    }
}
template <class T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    const int lane = (int)simt::cur->lane;
    const int lo = lane & ~(width - 1);
    const int s = (lane - (int)delta) < lo ? lane : lane - (int)delta;
    return __shfl_sync(mask, v, s, 32);
}
template <class T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    const int lane = (int)simt::cur->lane;
    const int hi = (lane & ~(width - 1)) + width - 1;
    const int s = (lane + (int)delta) > hi ? lane : lane + (int)delta;
    return __shfl_sync(mask, v, s, 32);
}
template <class T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    const int lane = (int)simt::cur->lane;
    int s = lane ^ lanemask;
    if ((s & ~(width - 1)) != (lane & ~(width - 1))) s = lane;
    return __shfl_sync(mask, v, s, 32);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) { return simt::collective(simt::OP_BALLOT, mask, pred ? 1u : 0u, 0); }
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
static inline unsigned __match_any_sync(unsigned mask, unsigned v) { return simt::collective(simt::OP_MATCH_ANY, mask, v, 0); }
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) { return simt::collective(simt::OP_MAX, mask, v, 0); }
static inline int __reduce_max_sync(unsigned mask, int v) { return (int)(simt::collective(simt::OP_MAX, mask, (uint32_t)v ^ 0x80000000u, 0) ^ 0x80000000u); }
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) { return simt::collective(simt::OP_MIN, mask, v, 0); }
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) { return simt::collective(simt::OP_ADD, mask, v, 0); }
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) { return simt::collective(simt::OP_OR, mask, v, 0); }
static inline unsigned __reduce_xor_sync(unsigned mask, unsigned v) { return simt::collective(simt::OP_XOR, mask, v, 0); }
static inline unsigned __reduce_and_sync(unsigned mask, unsigned v) { return simt::collective(simt::OP_AND, mask, v, 0); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { simt::collective(simt::OP_SYNC, mask, 0, 0); }
static inline void __syncthreads() { simt::syncthreads(); }
static inline void __threadfence_block() {}
static inline void __threadfence() {}
static inline void __nanosleep(unsigned) { simt::poll(); }
static inline unsigned __activemask() { return 0xffffffffu; }

// ---- integer intrinsics -------------------------------------------------------------------------
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
static inline int __clz(int x) { return __clz((unsigned)x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline unsigned __brev(unsigned x) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    const unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return (unsigned)(v >> (sh & 31u));
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) {
    const unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return (unsigned)((v << (sh & 31u)) >> 32);
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    const unsigned long long v = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned sel = (s >> (4 * i)) & 0xfu;
        unsigned byte = (unsigned)(v >> (8 * (sel & 7u))) & 0xffu;
        if (sel & 8u) byte = (byte & 0x80u) ? 0xffu : 0u;
        r |= byte << (8 * i);
    }
    return r;
}
static inline unsigned __vcmpeq2(unsigned a, unsigned b) {
    unsigned r = 0;
    if ((a & 0xffffu) == (b & 0xffffu)) r |= 0xffffu;
    if ((a >> 16) == (b >> 16)) r |= 0xffff0000u;
    return r;
}
static inline unsigned __vcmpeq4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int i = 0; i < 4; i++)
        if (((a >> (8 * i)) & 0xffu) == ((b >> (8 * i)) & 0xffu)) r |= 0xffu << (8 * i);
    return r;
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
template <class T>
static inline T __ldg(const T *p) { return *p; }
template <class T>
static inline T __ldcg(const T *p) { return *p; }
template <class T>
static inline T __ldcv(const T *p) { return *p; }
template <class T>
static inline void __stcg(T *p, T v) { *p = v; }
static inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)p; }
static inline long long clock64() { return simt::now(); }

// min / max as the CUDA headers provide them at global scope
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
static inline unsigned min(int a, unsigned b) { return (unsigned)a < b ? (unsigned)a : b; }
static inline unsigned max(unsigned a, int b) { return a > (unsigned)b ? a : (unsigned)b; }
static inline unsigned max(int a, unsigned b) { return (unsigned)a > b ? (unsigned)a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }

// ---- atomics (one OS thread, no pre-emption: plain read-modify-write) ---------------------------
template <class T, class U>
static inline T atomicAdd(T *p, U v) {
    const T old = *p;
    *p = (T)(old + (T)v);
    return old;
}
template <class T, class U>
static inline T atomicOr(T *p, U v) {
    const T old = *p;
    *p = (T)(old | (T)v);
    return old;
}
template <class T, class U>
static inline T atomicAnd(T *p, U v) {
    const T old = *p;
    *p = (T)(old & (T)v);
    return old;
}
template <class T, class U>
static inline T atomicMax(T *p, U v) {
    const T old = *p;
    if ((T)v > old) *p = (T)v;
    return old;
}
template <class T, class U>
static inline T atomicMin(T *p, U v) {
    const T old = *p;
    if ((T)v < old) *p = (T)v;
    return old;
}
template <class T, class U>
static inline T atomicExch(T *p, U v) {
    const T old = *p;
    *p = (T)v;
    return old;
}
template <class T, class U, class V>
static inline T atomicCAS(T *p, U cmp, V v) {
    const T old = *p;
    if (old == (T)cmp) *p = (T)v;
    return old;
}

// ---- the slice of the CUDA runtime API that capi.cu uses ---------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorLaunchFailure = 719 };
typedef void *cudaStream_t;
typedef struct simt_event {
    double t;
} *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributeNonPortableClusterSizeAllowed = 10 };
struct cudaDeviceProp {
    char name[256];
    int major, minor, multiProcessorCount;
};
static inline cudaError_t cudaGetDeviceCount(int *n) {
    *n = 1;
    return cudaSuccess;
}
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    memset(p, 0, sizeof *p);
    strcpy(p->name, "SIMT emulator (tests only)");
    p->major = 10;
    p->minor = 0;
    p->multiProcessorCount = 148;
    return cudaSuccess;
}
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) {
    *v = 148;
    return cudaSuccess;
}
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) {
    *d = 0;
    return cudaSuccess;
}
cudaError_t cudaGetLastError();
static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : (simt::last_error() ? simt::last_error() : "emulated failure"); }
static inline cudaError_t cudaMalloc(void **p, size_t n) {
    *p = aligned_alloc(256, (n + 255) & ~(size_t)255);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <class T>
static inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
static inline cudaError_t cudaFree(void *p) {
    free(p);
    return cudaSuccess;
}
template <class T>
static inline cudaError_t cudaMallocHost(T **p, size_t n) { return cudaMalloc((void **)p, n); }
static inline cudaError_t cudaFreeHost(void *p) { return cudaFree(p); }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) {
    memmove(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) {
    memmove(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = nullptr) {
    for (size_t i = 0; i < h; i++) memmove((char *)d + i * dp, (const char *)s + i * sp, w);
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) {
    memset(d, v, n);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) {
    *s = (void *)1;
    return cudaSuccess;
}
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e);
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

// ---- call-site check --------------------------------------------------------------------------------------------
// On the GPU every lane named in the mask must execute THE SAME warp-collective instruction; two lanes that reach
// `__shfl_sync` at different places of a divergent branch is undefined behaviour (in practice a hang).  The fibre model
// would happily pair them up, so every textual use of a collective gets a site number (__COUNTER__) and the runtime
// fails the launch if the lanes of one collective disagree about it.
#define __shfl_sync(...) (simt::cur_site = (__LINE__ * 4096 + (__COUNTER__ & 4095) + 1), __shfl_sync(__VA_ARGS__))
#define __shfl_up_sync(...) (simt::cur_site = (__LINE__ * 4096 + (__COUNTER__ & 4095) + 1), __shfl_up_sync(__VA_ARGS__))
#define __shfl_down_sync(...) (simt::cur_site = (__LINE__ * 4096 + (__COUNTER__ & 4095) + 1), __shfl_down_sync(__VA_ARGS__))
#define __shfl_xor_sync(...) (simt::cur_site = (__LINE__ * 4096 + (__COUNTER__ & 4095) + 1), __shfl_xor_sync(__VA_ARGS__))
#define __ballot_sync(...) (simt::cur_site = (__LINE__ * 4096 + (__COUNTER__ & 4095) + 1), __ballot_sync(__VA_ARGS__))
#define __any_sync(...) (simt::cur_site = (__LINE__ * 4096 + (__COUNTER__ & 4095) + 1), __any_sync(__VA_ARGS__))
#define __all_sync(...) (simt::cur_site = (__LINE__ * 4096 + (__COUNTER__ & 4095) + 1), __all_sync(__VA_ARGS__))
#define __match_any_sync(...) (simt::cur_site = (__LINE__ * 4096 + (__COUNTER__ & 4095) + 1), __match_any_sync(__VA_ARGS__))
#define __syncwarp(...) (simt::cur_site = (__LINE__ * 4096 + (__COUNTER__ & 4095) + 1), __syncwarp(__VA_ARGS__))
