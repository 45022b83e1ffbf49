// simt_emu.h — TEST INFRASTRUCTURE.  A small SIMT emulator so that the *real* kernel sources
// of consent_b200/csrc (written in plain CUDA C++) can be compiled with g++ and executed on a
// CPU-only box, where they are checked against the oracle before any GPU time is spent.
//
// It is NOT a product path and NOT a CPU fallback: nothing in consent_b200/ loads the library
// built with it (tests/emu/libconsent_emu.so), only `-m "not gpu"` tests do, and the name says
// what it is.  The product library libconsent_b200.so is nvcc-compiled sm_100a code only.
//
// Model: a thread block is a set of fibers (one per CUDA thread, own stack, hand-rolled
// x86-64 context switch) scheduled round-robin on one OS thread; a fiber runs until it reaches
// a collective (__syncthreads, __syncwarp, __shfl*_sync, __ballot_sync ...), which is a
// generation barrier that yields.  Blocks of a grid are distributed over a pool of OS threads,
// so global-memory atomics are real atomics.  Warp collectives require the full mask.
//
// Limits: no races are detected (the schedule is deterministic), no timing, no textures.
#pragma once
#ifndef CG_EMU
#error "simt_emu.h is only for -DCG_EMU builds"
#endif
#if !defined(__x86_64__)
#error "simt_emu.h needs x86-64"
#endif

#include <sys/mman.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

// ------------------------------------------------------------------ CUDA keywords
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }

namespace cg_emu {

extern "C" void cg_emu_switch(void** from_sp, void* to_sp);
#ifdef CG_EMU_IMPL
asm(R"(
.text
.globl cg_emu_switch
.type cg_emu_switch,@function
cg_emu_switch:
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
.size cg_emu_switch,.-cg_emu_switch
)");
#endif

struct WarpState {
    unsigned long long slot[32];
    unsigned arrived = 0, gen = 0;
};
struct Block;
struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    bool done = true;
    dim3 tidx;
    Block* blk = nullptr;
};
struct Block {
    dim3 bidx, bdim, gdim;
    unsigned nthreads = 0;
    unsigned bar_arrived = 0, bar_gen = 0;
    unsigned long long progress = 0;      // barrier completions (deadlock detection)
    std::vector<WarpState> warps;
    unsigned char* dyn_smem = nullptr;
    void* sched_sp = nullptr;
    const std::function<void()>* body = nullptr;
};

static const size_t kStackBytes = 192 * 1024;
static const unsigned kMaxThreads = 1024;
static const size_t kSmemBytes = 232 * 1024;

struct Worker {
    std::vector<Fiber> fibers;
    char* stacks = nullptr;
    unsigned char* smem = nullptr;
    Worker() {
        stacks = (char*)mmap(nullptr, kStackBytes * kMaxThreads, PROT_READ | PROT_WRITE,
                             MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (stacks == (char*)MAP_FAILED) { perror("mmap"); abort(); }
        smem = (unsigned char*)aligned_alloc(128, kSmemBytes);
        fibers.resize(kMaxThreads);
        for (unsigned i = 0; i < kMaxThreads; ++i) fibers[i].stack = stacks + (size_t)i * kStackBytes;
    }
    ~Worker() { munmap(stacks, kStackBytes * kMaxThreads); free(smem); }
};

#ifdef CG_EMU_IMPL
thread_local Fiber* cur = nullptr;
int g_threads = 0;
#else
extern thread_local Fiber* cur;
extern int g_threads;
#endif

inline void yield() { Fiber* f = cur; cg_emu_switch(&f->sp, f->blk->sched_sp); }

#ifdef CG_EMU_IMPL
extern "C" void cg_emu_entry() {
    Fiber* f = cur;
    (*f->blk->body)();
    f->done = true;
    for (;;) yield();
}
void run_block(Worker& w, Block& b) {
    for (unsigned t = 0; t < b.nthreads; ++t) {
        Fiber& f = w.fibers[t];
        f.done = false; f.blk = &b;
        f.tidx = dim3(t % b.bdim.x, (t / b.bdim.x) % b.bdim.y, t / (b.bdim.x * b.bdim.y));
        uintptr_t top = ((uintptr_t)f.stack + kStackBytes) & ~(uintptr_t)15;
        void** sp0 = (void**)(top - 8 * sizeof(void*));
        for (int i = 0; i < 6; ++i) sp0[i] = nullptr;
        sp0[6] = (void*)&cg_emu_entry;
        sp0[7] = nullptr;
        f.sp = sp0;
    }
    unsigned alive = b.nthreads;
    while (alive) {
        unsigned a = 0;
        const unsigned long long before = b.progress;
        for (unsigned t = 0; t < b.nthreads; ++t) {
            Fiber& f = w.fibers[t];
            if (f.done) continue;
            cur = &f;
            cg_emu_switch(&b.sched_sp, f.sp);
            if (!f.done) ++a;
        }
        // fibers only yield inside barrier waits: a round with no barrier completed and nobody done is a deadlock
        if (a == alive && b.progress == before) { fprintf(stderr, "simt_emu: deadlock (block %u): a collective is not reached by all threads\n", b.bidx.x); abort(); }
        alive = a;
    }
    cur = nullptr;
}
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    unsigned nblocks = grid.x * grid.y * grid.z;
    unsigned nthreads = block.x * block.y * block.z;
    if (nthreads == 0 || nthreads > kMaxThreads || smem > kSmemBytes) { fprintf(stderr, "simt_emu: bad launch config\n"); abort(); }
    if (nblocks == 0) return;
    int nthr = g_threads > 0 ? g_threads : (int)std::thread::hardware_concurrency();
    if (const char* e = getenv("CG_EMU_THREADS")) nthr = atoi(e);
    if (nthr < 1) nthr = 1;
    if ((unsigned)nthr > nblocks) nthr = (int)nblocks;
    std::atomic<unsigned> next(0);
    auto work = [&]() {
        Worker* w = new Worker();
        for (;;) {
            unsigned bid = next.fetch_add(1);
            if (bid >= nblocks) break;
            Block b;
            b.bidx = dim3(bid % grid.x, (bid / grid.x) % grid.y, bid / (grid.x * grid.y));
            b.bdim = block; b.gdim = grid; b.nthreads = nthreads;
            b.warps.resize((nthreads + 31) / 32);
            b.dyn_smem = w->smem;
            memset(w->smem, 0xCD, smem);          // poison: shared memory is uninitialised on a GPU
            b.body = &body;
            run_block(*w, b);
        }
        delete w;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthr; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
}
#else
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
#endif

inline void block_barrier() {
    Block* b = cur->blk;
    unsigned g = b->bar_gen;
    if (++b->bar_arrived == b->nthreads) { b->bar_arrived = 0; b->bar_gen = g + 1; b->progress++; return; }
    while (b->bar_gen == g) yield();
}
inline unsigned linear_tid() { Fiber* f = cur; return f->tidx.x + f->blk->bdim.x * (f->tidx.y + f->blk->bdim.y * f->tidx.z); }
inline WarpState& my_warp() { return cur->blk->warps[linear_tid() / 32]; }
inline unsigned warp_size_here() {
    Block* b = cur->blk; unsigned w = linear_tid() / 32;
    unsigned rest = b->nthreads - w * 32; return rest < 32 ? rest : 32;
}
inline void warp_barrier() {
    WarpState& w = my_warp();
    unsigned n = warp_size_here();
    unsigned g = w.gen;
    if (++w.arrived == n) { w.arrived = 0; w.gen = g + 1; cur->blk->progress++; return; }
    while (w.gen == g) yield();
}
inline void check_mask(unsigned mask) {
    if (mask != 0xffffffffu) { fprintf(stderr, "simt_emu: partial warp masks are not supported\n"); abort(); }
}
template <class T> inline T warp_exchange(T v, unsigned src_lane) {
    static_assert(sizeof(T) <= 8, "shfl payload");
    WarpState& w = my_warp();
    unsigned lane = linear_tid() & 31;
    unsigned long long raw = 0; memcpy(&raw, &v, sizeof(T));
    w.slot[lane] = raw;
    warp_barrier();
    unsigned long long got = w.slot[src_lane & 31];
    warp_barrier();
    T r; memcpy(&r, &got, sizeof(T));
    return r;
}
}  // namespace cg_emu

#define threadIdx (cg_emu::cur->tidx)
#define blockIdx  (cg_emu::cur->blk->bidx)
#define blockDim  (cg_emu::cur->blk->bdim)
#define gridDim   (cg_emu::cur->blk->gdim)
#define warpSize  32

static inline void __syncthreads() { cg_emu::block_barrier(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { cg_emu::check_mask(mask); cg_emu::warp_barrier(); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }

template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    cg_emu::check_mask(mask);
    unsigned lane = cg_emu::linear_tid() & 31;
    unsigned base = lane & ~(unsigned)(width - 1);
    return cg_emu::warp_exchange(v, base + ((unsigned)src & (unsigned)(width - 1)));
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
    cg_emu::check_mask(mask);
    unsigned lane = cg_emu::linear_tid() & 31;
    unsigned base = lane & ~(unsigned)(width - 1);
    unsigned src = (lane - base >= d) ? lane - d : lane;
    return cg_emu::warp_exchange(v, src);
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
    cg_emu::check_mask(mask);
    unsigned lane = cg_emu::linear_tid() & 31;
    unsigned base = lane & ~(unsigned)(width - 1);
    unsigned src = (lane - base + d < (unsigned)width) ? lane + d : lane;
    return cg_emu::warp_exchange(v, src);
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32) {
    cg_emu::check_mask(mask);
    unsigned lane = cg_emu::linear_tid() & 31;
    return cg_emu::warp_exchange(v, lane ^ (unsigned)x);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    cg_emu::check_mask(mask);
    cg_emu::WarpState& w = cg_emu::my_warp();
    unsigned lane = cg_emu::linear_tid() & 31;
    w.slot[lane] = pred ? 1 : 0;
    cg_emu::warp_barrier();
    unsigned r = 0, n = cg_emu::warp_size_here();
    for (unsigned i = 0; i < n; ++i) if (w.slot[i]) r |= 1u << i;
    cg_emu::warp_barrier();
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __reduce_max_sync(unsigned mask, int v) {
    for (int d = 16; d > 0; d >>= 1) { const int o = __shfl_xor_sync(mask, v, d); v = o > v ? o : v; }
    return v;
}
static inline int __all_sync(unsigned mask, int pred) {
    unsigned n = cg_emu::warp_size_here();
    unsigned full = n == 32 ? 0xffffffffu : ((1u << n) - 1);
    return __ballot_sync(mask, pred) == full;
}

// ------------------------------------------------------------------ integer intrinsics
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; ++i) if (x & (1u << i)) r |= 1u << (31 - i); return r; }
static inline unsigned __vadd2(unsigned a, unsigned b) {          // per-halfword add, wrap-around
    return (((a & 0xffffu) + (b & 0xffffu)) & 0xffffu) | ((((a >> 16) + (b >> 16)) & 0xffffu) << 16);
}
static inline unsigned __vmaxs2(unsigned a, unsigned b) {         // per-halfword signed maximum
    const short al = (short)(a & 0xffffu), bl = (short)(b & 0xffffu), ah = (short)(a >> 16), bh = (short)(b >> 16);
    return (unsigned)(unsigned short)(al > bl ? al : bl) | ((unsigned)(unsigned short)(ah > bh ? ah : bh) << 16);
}
static inline unsigned __viaddmax_s16x2(unsigned a, unsigned b, unsigned c) { return __vmaxs2(__vadd2(a, b), c); }            // per halfword max(a + b, c)
static inline unsigned __viaddmax_s16x2_relu(unsigned a, unsigned b, unsigned c) { return __vmaxs2(__vmaxs2(__vadd2(a, b), c), 0u); }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }                                      // one rounding each, no contraction
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned sel) {
    const unsigned long long v = ((unsigned long long)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) r |= (unsigned)((v >> (8 * ((sel >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return r;
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift) {
    unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return (unsigned)((v << (shift & 31)) >> 32);
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {
    unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return (unsigned)(v >> (shift & 31));
}
template <class T> static inline T __ldg(const T* p) { return *p; }
using std::max;
using std::min;

// ------------------------------------------------------------------ atomics
template <class T> static inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicSub(T* p, T v) { return __atomic_fetch_sub(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicAnd(T* p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicMax(T* p, T v) {
    T old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
template <class T> static inline T atomicMin(T* p, T v) {
    T old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
template <class T> static inline T atomicCAS(T* p, T cmp, T v) {
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
    return cmp;
}

// ------------------------------------------------------------------ runtime shim
typedef int cudaError_t;
typedef struct cg_emu_stream_* cudaStream_t;
struct cg_emu_event_ { std::chrono::steady_clock::time_point t; };
typedef cg_emu_event_* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNoDevice = 100 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
enum { cudaStreamNonBlocking = 1, cudaEventDefault = 0, cudaEventDisableTiming = 2 };
static inline const char* cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "emulated error"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
    *v = (a == cudaDevAttrMultiProcessorCount) ? 4 : (int)cg_emu::kSmemBytes - 1024; return 0;
}
static inline cudaError_t cudaMalloc(void** p, size_t n) {
    *p = aligned_alloc(256, (n + 255) & ~(size_t)255);
    if (!*p) return cudaErrorMemoryAllocation;
    memset(*p, 0xA5, n);                          // poison: cudaMalloc memory is uninitialised
    return 0;
}
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) & ~(size_t)255); return *p ? 0 : cudaErrorMemoryAllocation; }
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMallocHost((void**)p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = nullptr; return 0; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = 0; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new cg_emu_event_(); return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new cg_emu_event_(); return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return 0;
}
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }
static inline cudaError_t cudaMemGetInfo(size_t* fr, size_t* tot) { *fr = (size_t)8 << 30; *tot = (size_t)8 << 30; return 0; }

#define CG_LAUNCH(kernel, grid, block, smem, stream, ...) \
    cg_emu::launch(dim3(grid), dim3(block), (smem), [=]() { kernel(__VA_ARGS__); })
#define CG_DYN_SMEM(name) unsigned char* name = cg_emu::cur->blk->dyn_smem
// the sm_100a-only helpers of cg_common.cuh: one count table per CTA instead of per SM, no prefetch, the bulk copies as plain copies
static inline unsigned cg_smid() { return blockIdx.x; }
static inline void cg_prefetch_l1(const void*) {}
static inline void cg_bulk_issue2(void* dst0, const void* src0, void* dst1, const void* src1, unsigned bytes, uint64_t*) { memcpy(dst0, src0, bytes); memcpy(dst1, src1, bytes); }
static inline void cg_bulk_wait(uint64_t*) {}
