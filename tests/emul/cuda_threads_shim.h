// cuda_threads_shim.h -- TEST INFRASTRUCTURE: a host model of the slice of CUDA the
// K2 counting-sort kernel uses, with ONE OS THREAD PER CUDA THREAD:
//   * threadIdx / blockIdx / gridDim, dynamic shared memory (exact size, 1024-byte
//     aligned, garbage-filled), __syncthreads;
//   * warp intrinsics (__shfl*_sync, __ballot_sync, __any_sync, __reduce_*_sync,
//     __syncwarp) as a 32-thread barrier around an exchange buffer -- there is NO
//     implicit lockstep: code that relies on it without a *_sync shows up as a
//     data race under ThreadSanitizer;
//   * shared / global atomics and fences as C++ atomics and fences;
//   * the mbarrier (phase counter + pending transaction bytes in its 64-bit word)
//     and the 2-D TMA tile load (box copy with the 128-byte swizzle, out-of-bounds
//     elements = NaN, completion on the mbarrier), performed either at issue time
//     (the earliest the hardware may write) or by a delayed helper thread.
// Included by tests/emul/k2_csort_threads.cpp before the kernel header.
#pragma once
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
#include <vector>

struct ShimDim { unsigned x; };
static thread_local ShimDim threadIdx, blockIdx, gridDim, blockDim;

struct ShimWarp {
    pthread_barrier_t bar;
    uint32_t xch[32];
};
struct ShimBlock {
    pthread_barrier_t bar;
    unsigned char* smem;
    size_t smem_bytes;
    ShimWarp* warps;
    int tma_mode;                       // 0: copy at issue, 1: delayed helper thread
    std::mutex mu;
    std::vector<std::thread> helpers;
};
static thread_local ShimBlock* g_blk;
static thread_local ShimWarp* g_warp;
static thread_local int g_lane;

static inline unsigned char* gsb_host_smem() { return g_blk->smem; }
static inline uint32_t gsb_smem_u32(const void* p) {
    return static_cast<uint32_t>(reinterpret_cast<const unsigned char*>(p) - g_blk->smem);
}
static inline void* gsb_host_shared_ptr(uint32_t addr) {
    if (addr + 4 > g_blk->smem_bytes) abort();
    return g_blk->smem + addr;
}

static inline void __syncthreads() { pthread_barrier_wait(&g_blk->bar); }
static inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&g_warp->bar); }

template <class T>
static inline T shim_exchange(T v, int src) {
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "32- or 64-bit values only");
    uint32_t b[2] = {0u, 0u};
    memcpy(b, &v, sizeof(T));
    uint32_t r[2];
    for (int h = 0; h < static_cast<int>(sizeof(T) / 4); ++h) {
        g_warp->xch[g_lane] = b[h];
        pthread_barrier_wait(&g_warp->bar);
        r[h] = g_warp->xch[src & 31];
        pthread_barrier_wait(&g_warp->bar);
    }
    T out;
    memcpy(&out, r, sizeof(T));
    return out;
}
template <class T> static inline T __shfl_sync(unsigned m, T v, int src) { if (m != 0xffffffffu) abort(); return shim_exchange(v, src); }
template <class T> static inline T __shfl_xor_sync(unsigned m, T v, int x) { if (m != 0xffffffffu) abort(); return shim_exchange(v, g_lane ^ x); }
template <class T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d) {
    if (m != 0xffffffffu) abort();
    const int src = g_lane - static_cast<int>(d);
    return shim_exchange(v, src < 0 ? g_lane : src);
}
static inline unsigned __ballot_sync(unsigned m, int pred) {
    if (m != 0xffffffffu) abort();
    g_warp->xch[g_lane] = pred ? 1u : 0u;
    pthread_barrier_wait(&g_warp->bar);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= g_warp->xch[i] << i;
    pthread_barrier_wait(&g_warp->bar);
    return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
static inline unsigned __reduce_min_sync(unsigned m, unsigned v) {
    if (m != 0xffffffffu) abort();
    g_warp->xch[g_lane] = v;
    pthread_barrier_wait(&g_warp->bar);
    unsigned r = 0xffffffffu;
    for (int i = 0; i < 32; ++i) r = g_warp->xch[i] < r ? g_warp->xch[i] : r;
    pthread_barrier_wait(&g_warp->bar);
    return r;
}
static inline unsigned __reduce_max_sync(unsigned m, unsigned v) {
    if (m != 0xffffffffu) abort();
    g_warp->xch[g_lane] = v;
    pthread_barrier_wait(&g_warp->bar);
    unsigned r = 0u;
    for (int i = 0; i < 32; ++i) r = g_warp->xch[i] > r ? g_warp->xch[i] : r;
    pthread_barrier_wait(&g_warp->bar);
    return r;
}

// CUDA atomics are relaxed and the kernel orders them with __threadfence_block;
// ThreadSanitizer does not model stand-alone fences, so the ordering the fence +
// atomic pair gives on the device is put on the atomic itself (acquire-release)
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL);
}
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline long long clock64() { return 0; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs(static_cast<int>(v)); }
// a NaN PRODUCED by GPU arithmetic is the canonical 0x7fffffff; x86 produces the
// negative default NaN (INF * 0 in the bin map of an empty column): same bits here
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return f != f ? 0x7fffffffu : u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return f != f ? 0x7fffffff : u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __viaddmin_s32_relu(int a, int b, int c) {
    long long s = static_cast<long long>(a) + b;
    if (s > c) s = c;
    if (s < 0) s = 0;
    return static_cast<int>(s);
}
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
template <class T> static inline T max(T a, T b) { return a > b ? a : b; }
template <class T> static inline T min(T a, T b) { return a < b ? a : b; }

// ---- mbarrier: low 32 bits = completed phases, high 32 bits = pending bytes
static inline void gsb_mbar_init(uint64_t* bar, uint32_t) { __atomic_store_n(bar, 0ull, __ATOMIC_SEQ_CST); }
static inline void gsb_fence_barrier_init() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void gsb_fence_proxy_async() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void gsb_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    __atomic_fetch_add(bar, static_cast<uint64_t>(bytes) << 32, __ATOMIC_ACQ_REL);
}
static inline void shim_mbar_complete_tx(uint64_t* bar, uint32_t bytes) {
    uint64_t o = __atomic_load_n(bar, __ATOMIC_RELAXED);
    for (;;) {
        if ((o >> 32) < bytes) abort();                       // more bytes than expected
        uint64_t n = o - (static_cast<uint64_t>(bytes) << 32);
        if ((n >> 32) == 0) n = (n & 0xffffffff00000000ull) | ((n + 1) & 0xffffffffull);
        if (__atomic_compare_exchange_n(bar, &o, n, false, __ATOMIC_ACQ_REL, __ATOMIC_RELAXED)) return;
    }
}
static inline void gsb_mbar_wait(uint64_t* bar, uint32_t parity) {
    long spins = 0;
    while (((__atomic_load_n(bar, __ATOMIC_ACQUIRE) & 0xffffffffull) & 1ull) == parity) {
        sched_yield();
        if (++spins > 200000000L) abort();                    // a deadlock is a failure, not a hang
    }
}

// ---- 2-D TMA tile load: the descriptor the emulation stores in the CUtensorMap
struct ShimTmap {
    const float* data;      // [R][pitch]
    long long pitch, N;
    int R, box_rows;
};
static inline void shim_tma_copy(unsigned char* dst, const ShimTmap& d, int c0, int r0) {
    const uint32_t nanb = 0x7fc00000u;
    for (int r = 0; r < d.box_rows; ++r)
        for (int c = 0; c < 32; ++c) {
            const long long col = static_cast<long long>(c0) + c;
            const int row = r0 + r;
            uint32_t v = nanb;
            if (row < d.R && col < d.N && col >= 0) memcpy(&v, d.data + static_cast<long long>(row) * d.pitch + col, 4);
            // 128-byte swizzle: 16-byte chunk index ^ (row & 7) (rows are 128 bytes)
            const uint32_t off = static_cast<uint32_t>(r) * 128u + ((((static_cast<uint32_t>(c) >> 2) ^ (static_cast<uint32_t>(row) & 7u)) << 4)) + (c & 3) * 4u;
            memcpy(dst + off, &v, 4);
        }
}
static inline void gsb_tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    ShimTmap d;
    memcpy(&d, map, sizeof(d));
    unsigned char* dst = static_cast<unsigned char*>(smem_dst);
    const uint32_t bytes = static_cast<uint32_t>(d.box_rows) * 128u;
    if (g_blk->tma_mode == 0) {
        shim_tma_copy(dst, d, c0, c1);
        shim_mbar_complete_tx(bar, bytes);
    } else {
        std::lock_guard<std::mutex> lk(g_blk->mu);
        g_blk->helpers.emplace_back([dst, d, c0, c1, bar, bytes]() {
            std::this_thread::sleep_for(std::chrono::microseconds(300));
            shim_tma_copy(dst, d, c0, c1);
            shim_mbar_complete_tx(bar, bytes);
        });
    }
}
static inline void gsb_prefetch_tmap(const CUtensorMap*) {}

// ---- launching a block: NT threads, warps of 32
template <class F>
static void shim_run_block(int block, int grid, int nthreads, size_t smem_bytes, int tma_mode, F kernel) {
    ShimBlock blk;
    pthread_barrier_init(&blk.bar, nullptr, nthreads);
    const int nwarps = nthreads / 32;
    std::vector<ShimWarp> warps(nwarps);
    for (auto& w : warps) pthread_barrier_init(&w.bar, nullptr, 32);
    void* mem = nullptr;
    if (posix_memalign(&mem, 1024, smem_bytes)) abort();
    memset(mem, 0xA5, smem_bytes);
    blk.smem = static_cast<unsigned char*>(mem);
    blk.smem_bytes = smem_bytes;
    blk.warps = warps.data();
    blk.tma_mode = tma_mode;
    std::vector<std::thread> th;
    th.reserve(nthreads);
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([&blk, block, grid, t, nthreads, &kernel]() {
            g_blk = &blk;
            g_warp = blk.warps + (t >> 5);
            g_lane = t & 31;
            threadIdx.x = static_cast<unsigned>(t);
            blockIdx.x = static_cast<unsigned>(block);
            gridDim.x = static_cast<unsigned>(grid);
            blockDim.x = static_cast<unsigned>(nthreads);
            kernel();
        });
    for (auto& t : th) t.join();
    for (auto& h : blk.helpers) h.join();
    for (auto& w : warps) pthread_barrier_destroy(&w.bar);
    pthread_barrier_destroy(&blk.bar);
    free(mem);
}
