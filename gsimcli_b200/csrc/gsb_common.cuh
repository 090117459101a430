// gsb_common.cuh -- shared declarations of libgsimcli_b200 (sm_100a only)
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <string.h>
#include "../../include/gsimcli_b200.h"

// ------------------------------------------------------------------ errors
void gsb_set_error(const char* fmt, ...);

#define GSB_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t e_ = (call);                                              \
        if (e_ != cudaSuccess) {                                              \
            gsb_set_error("%s failed: %s (%s:%d)", #call,                     \
                          cudaGetErrorString(e_), __FILE__, __LINE__);        \
            return (e_ == cudaErrorMemoryAllocation) ? GSB_ENOMEM : GSB_ECUDA;\
        }                                                                     \
    } while (0)

#define GSB_REQUIRE(cond, ...)                                                \
    do {                                                                      \
        if (!(cond)) {                                                        \
            gsb_set_error(__VA_ARGS__);                                       \
            return GSB_EINVAL;                                                \
        }                                                                     \
    } while (0)

// ------------------------------------------------------------------- stack
// [R][pitch] float32 in HBM, realization-major.  pitch is a multiple of 32
// floats so every row starts on a 128-byte line and the TMA row stride is a
// multiple of 16 bytes.
struct gsb_stack {
    int device;
    int sm_count;
    int64_t R;
    int64_t N;
    int64_t pitch;
    float* data;
    // device buffers reused across calls (grown on demand, freed with the
    // stack): `scratch` belongs to the kernel launchers, `stage` to the C-ABI
    // wrappers (host<->device staging of text, node ids, tables, maps)
    void* scratch;
    size_t scratch_bytes;
    void* stage;
    size_t stage_bytes;
    void* pinned;
    size_t pinned_bytes;
    cudaEvent_t ev0, ev1;
    CUtensorMap tmap_k2;   // 2-D map (node, realization): box 32 nodes x
                           // gsb_k2_box_rows(R) rows, 128B swizzle, OOB -> NaN
    int tmap_ok;
    int k2_sm_reserve;     // SMs the K2 grid leaves free (gsb_stack_set_option)
    int k2_algo;           // tests only: 0 = default, 1 = counting sort, 2 = bitonic, 3 = register tile, 4 = staged select
    int k2_nw;             // tests only: warps per CTA of the register-tile kernel (0 = default)
    int k2_stats;          // diagnostic counters of the counting-sort kernel
    int k1_algo;           // tests only: 1 = the experimental 8-records-per-thread DSS decoder
    int* k2_flags;         // device, [0] count + one slot per 32-node tile: tiles the register-
                           // tile kernel hands to the counting-sort kernel (own buffer: K2 may
                           // run concurrently with calls that regrow `scratch` on other streams)
};

int gsb_scratch(gsb_stack* s, size_t bytes, void** out);
int gsb_stage(gsb_stack* s, size_t bytes, void** out);
int gsb_pinned(gsb_stack* s, size_t bytes, void** out);

// ------------------------------------------------------- output plane kinds
enum {
    GSB_PK_MEAN = 0,
    GSB_PK_QUANT = 1,   // quantile q (median is q = 0.5)
    GSB_PK_SKEW = 2,
    GSB_PK_VAR = 3,
    GSB_PK_STD = 4,
    GSB_PK_COEFVAR = 5,
    GSB_PK_MODE = 6
};

struct GsbPlanes {
    int n;
    int need_sort;     // any quantile or mode plane
    int need_mode;
    unsigned char kind[GSB_MAX_PLANES];
    double q[GSB_MAX_PLANES];
};

// builds the plane list in the documented order; returns GSB_EINVAL when it
// does not fit
int gsb_build_planes(uint32_t flags, double p, const double* q, int nq,
                     GsbPlanes* out);

int gsb_k2_box_rows(int64_t R);   // rows per TMA box of the K2 tile (k2_grid_sort.cu)

// ---------------------------------------------------------------- launchers
int gsb_launch_k2_sort(gsb_stack* s, const GsbPlanes& pl, float nodata,
                       int64_t n0, int64_t nn, float* d_out, int64_t out_pitch,
                       cudaStream_t st);
int gsb_launch_k2_moments(gsb_stack* s, const GsbPlanes& pl, float nodata,
                          int64_t n0, int64_t nn, float* d_out,
                          int64_t out_pitch, cudaStream_t st);
int gsb_launch_k2_deep(gsb_stack* s, const GsbPlanes& pl, float nodata, int64_t n0, int64_t nn,
                       float* d_out, int64_t out_pitch, cudaStream_t st);
int gsb_launch_k2_columns(gsb_stack* s, const int64_t* d_node_ids, int S,
                          int dz, int K, const GsbPlanes& pl, float nodata,
                          double* d_table, cudaStream_t st);
int gsb_launch_extract(gsb_stack* s, const int64_t* d_node_ids, int64_t ncols,
                       int K, float* d_values, cudaStream_t st);
int gsb_launch_k3_detect(const double* clim, const double* mean,
                         const double* lperc, const double* rperc,
                         const double* fix_a, const double* fix_b,
                         const double* fix_c, int S, int dz, int method,
                         double thr, double nodata, double* filled,
                         double* homog, double* flag, int32_t* detected,
                         cudaStream_t st);
int gsb_launch_k3_detect_table(const double* clim, const double* table, const double* fixtab,
                               const int64_t* zoff, int64_t sstride, int C, int Cfix, int c_mean,
                               int c_lperc, int c_rperc, int c_fa, int c_fb, int c_fc, int S, int dz,
                               int method, double thr, double nodata, double* out,
                               int32_t* detected, cudaStream_t st);
int gsb_launch_synth(gsb_stack* s, uint64_t seed, int64_t dx, int64_t dy,
                     int64_t dz, int64_t node0, float nodata, cudaStream_t st);
int gsb_launch_decode(gsb_stack* s, int64_t r, const char* dtext, size_t nbytes,
                      int header_lines, int line_width, int64_t first_line,
                      int64_t n0, int64_t nlines, int64_t* d_bad_line,
                      cudaStream_t st);
int gsb_launch_encode(const gsb_stack* s, int64_t r, int64_t n0, int64_t nn,
                      int width, int prec, int crlf, char* dtext,
                      cudaStream_t st);

// ------------------------------------------------------- device primitives
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t gsb_smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void gsb_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gsb_smem_u32(bar)),
                 "r"(count));
}
__device__ __forceinline__ void gsb_fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void gsb_fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void gsb_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gsb_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void gsb_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                     gsb_smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void gsb_mbar_wait(uint64_t* bar, uint32_t parity) {
    // try_wait suspends the thread in hardware up to the time hint (ns) before it
    // reports "not yet": a long hint keeps waiting warps out of the issue slots
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(gsb_smem_u32(bar)),
        "r"(parity), "r"(1000000u)
        : "memory");
}
// 2-D TMA tile load, global -> shared::cta, completion on an mbarrier
__device__ __forceinline__ void gsb_tma_load_2d(void* smem_dst, const CUtensorMap* map,
                                                uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(gsb_smem_u32(smem_dst)),
        "l"(map), "r"(gsb_smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void gsb_prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
#endif
