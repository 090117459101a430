// k2_ssel_threads.cpp -- TEST INFRASTRUCTURE: the tile loop of the K2 staged-select
// kernel (gsimcli_b200/csrc/k2_ssel_kernel.cuh: the kernel function itself, with
// its barriers, its cp.async staging of the next tile, its shared-memory layout
// and its output stores) compiled for the host and run with ONE OS THREAD PER
// CUDA THREAD and a real barrier.  What the serial emulation (k2_ssel_emul.cpp)
// cannot see shows up here as a data race under ThreadSanitizer or as a wrong
// plane: a missing barrier, a stage refilled while a warp still reads it, a
// region aliased too early, a bad offset of the layout (the shared block has its
// exact size: AddressSanitizer sees an overrun).
// The asynchronous copy is emulated at both ends of its freedom: mode 0 performs
// it at issue time (the earliest the hardware may write the stage), mode 1 at
// cp.async.wait_group (the latest).  Never linked into libgsimcli_b200.so.
//   g++ -O1 -g -std=c++17 -pthread -ffp-contract=off [-fsanitize=thread] -shared -fPIC ...
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <thread>
#include <utility>
#include <vector>

#define SSEL_HOST 1

// ---- the slice of the CUDA programming model the kernel uses
struct Dim1 { unsigned x; };
static thread_local Dim1 threadIdx, blockIdx, gridDim;
struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

struct BlockCtx {
    pthread_barrier_t bar;
    std::atomic<int> orflag;
    unsigned char* smem;
    size_t smem_bytes;
    int copy_mode;
};
static thread_local BlockCtx* g_ctx;
static thread_local std::vector<std::pair<float*, const float*>> g_pending;

static inline void __syncthreads() { pthread_barrier_wait(&g_ctx->bar); }
static inline int __syncthreads_or(int p) {
    BlockCtx* c = g_ctx;
    if (p) c->orflag.store(1, std::memory_order_relaxed);
    pthread_barrier_wait(&c->bar);
    const int r = c->orflag.load(std::memory_order_relaxed);
    pthread_barrier_wait(&c->bar);
    if (threadIdx.x == 0) c->orflag.store(0, std::memory_order_relaxed);   // next use is barriers away
    return r;
}
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned char* ssel_host_smem(size_t bytes) {
    if (bytes != g_ctx->smem_bytes) abort();
    return g_ctx->smem;
}
namespace ssel {
static inline void cp_async16(float* dst, const float* src) {
    if (g_ctx->copy_mode == 0) memcpy(dst, src, 16);
    else g_pending.emplace_back(dst, src);
}
static inline void cp_async_commit() {}
static inline void cp_async_wait_all() {
    for (auto& pr : g_pending) memcpy(pr.first, pr.second, 16);
    g_pending.clear();
}
}  // namespace ssel

#include "../../gsimcli_b200/csrc/k2_ssel_kernel.cuh"

using namespace ssel;

template <int NW, int IPT>
static void run_grid(const SselLaunch& L, int grid, int copy_mode) {
    constexpr int NT = NW * 32;
    constexpr size_t BYTES = Layout<NW, IPT>::BYTES;
    for (int b = 0; b < grid; ++b) {
        BlockCtx ctx;
        pthread_barrier_init(&ctx.bar, nullptr, NT);
        ctx.orflag.store(0);
        // exact size, 128-byte aligned, garbage-filled: nothing may rely on zeroed shared memory
        void* mem = nullptr;
        if (posix_memalign(&mem, 128, BYTES)) abort();
        memset(mem, 0xA5, BYTES);
        ctx.smem = static_cast<unsigned char*>(mem);
        ctx.smem_bytes = BYTES;
        ctx.copy_mode = copy_mode;
        std::vector<std::thread> th;
        th.reserve(NT);
        for (int t = 0; t < NT; ++t)
            th.emplace_back([&ctx, &L, b, grid, t]() {
                g_ctx = &ctx;
                threadIdx.x = static_cast<unsigned>(t);
                blockIdx.x = static_cast<unsigned>(b);
                gridDim.x = static_cast<unsigned>(grid);
                g_pending.clear();
                k2_ssel_kernel<NW, IPT>(L);
            });
        for (auto& t : th) t.join();
        pthread_barrier_destroy(&ctx.bar);
        free(mem);
    }
}

// Same argument meaning as gsb_launch_k2_sort + gsb_launch_k2_ssel: the stack is
// [R][pitch] (pitch a multiple of 32), the requested columns are [n0, n0 + nn),
// out is [nplanes][out_pitch] with column n0 at index 0.  flagged: [0] count,
// [1 + k] tile.  Returns 0, or -1 for a shape the kernel does not serve.
extern "C" int k2_ssel_threads(const float* data, long long pitch, int R, float nodata, long long n0,
                               long long nn, int nplanes, const unsigned char* kind, const double* q,
                               float* out, long long out_pitch, int nw, int grid, int copy_mode,
                               int* flagged) {
    if (R < 1 || R > 1024 || pitch < 32 || (pitch & 31)) return -1;
    SselLaunch L;
    memset(&L, 0, sizeof(L));
    L.a.R = R;
    L.a.nplanes = nplanes;
    L.a.nodata = (nodata == nodata) ? nodata : INFINITY;
    int nq = 0;
    for (int i = 0; i < nplanes; ++i) {
        L.a.kind[i] = kind[i];
        L.a.q[i] = q[i];
        if (kind[i] == PK_QUANT) L.a.qplane[nq++] = (unsigned char)i;
    }
    L.a.nq = nq;
    if (nq < 1 || 2 * nq > NLMAX) return -1;
    const long long head = n0 & 3;
    L.data = data;
    L.pitch = pitch;
    L.n0 = n0 - head;
    L.nn = nn + head;
    L.skip = head;
    L.ntiles = (L.nn + 31) / 32;
    L.out = out - head;
    L.out_pitch = out_pitch;
    L.flagged = flagged;
    flagged[0] = 0;
    if (grid > L.ntiles) grid = (int)L.ntiles;
    if (grid < 1) return 0;
    if (nw == 32) {
        if (R <= 256) run_grid<32, 8>(L, grid, copy_mode);
        else if (R <= 512) run_grid<32, 16>(L, grid, copy_mode);
        else run_grid<32, 32>(L, grid, copy_mode);
    } else {
        if (R <= 256) run_grid<16, 16>(L, grid, copy_mode);
        else if (R <= 512) run_grid<16, 32>(L, grid, copy_mode);
        else run_grid<16, 64>(L, grid, copy_mode);
    }
    return 0;
}
