// k2_columns.cu -- K2c: station columns (GridFiles.stats_area) and K3 detect
//
// K2c restates the layer loop of GridFiles.stats_area
// (GSIMCLI/tools/grid.py:812-850): for each station and time layer, the R*K
// values of the disc neighbourhood are gathered in the reference's order
// arr[i + j*K] (grid.py:823), nodata -> dropped (grid.py:829) and reduced in
// float64.  One CTA per (station, layer); the gather is a strided read of at
// most a few thousand values, negligible next to the full-grid kernel, so the
// design goal here is exactness: order statistics are selected float32
// elements of a block-wide bitonic sort, the quantile interpolation follows
// NumPy's _lerp operation by operation (no FMA contraction), sums are
// fixed-order float64 reductions.
//
// K3 restates detect()'s arithmetic (GSIMCLI/tools/homog.py:191-239): fill
// missing observations with the local mean, flag values outside the inclusive
// [lperc, rperc] interval, substitute the correction value, count detections.
#include "k2_columns_kernel.cuh"

namespace {
using namespace kcol;
}  // namespace

int gsb_launch_k2_columns(gsb_stack* s, const int64_t* d_node_ids, int S, int dz, int K,
                          const GsbPlanes& pl, float nodata, double* d_table,
                          cudaStream_t st) {
    const long long ncols = static_cast<long long>(S) * dz;
    if (ncols <= 0 || pl.n <= 0) return GSB_OK;
    const long long total = static_cast<long long>(s->R) * K;
    GSB_REQUIRE(total <= (1 << 22), "stats_columns: R*K = %lld too large", total);
    int npow2 = 32;
    while (npow2 < total) npow2 <<= 1;
    KcArgs a;
    memset(&a, 0, sizeof(a));
    a.data = s->data;
    a.pitch = s->pitch;
    a.R = static_cast<int>(s->R);
    a.K = K;
    a.nplanes = pl.n;
    a.npow2 = npow2;
    a.nodata = nodata;
    a.node_ids = reinterpret_cast<const long long*>(d_node_ids);
    a.table = d_table;
    for (int i = 0; i < pl.n; ++i) { a.kind[i] = pl.kind[i]; a.q[i] = pl.q[i]; }
    size_t smem = static_cast<size_t>(npow2) * 4;
    a.use_global = smem > 200 * 1024;
    if (a.use_global) {
        void* g = nullptr;
        int rc = gsb_scratch(s, static_cast<size_t>(ncols) * npow2 * 4, &g);
        if (rc) return rc;
        a.gsort = static_cast<float*>(g);
        smem = 0;
    }
    GSB_CUDA(cudaFuncSetAttribute(k2_columns_kernel,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  200 * 1024));
    k2_columns_kernel<<<static_cast<int>(ncols), KC_THREADS, smem, st>>>(a);
    GSB_CUDA(cudaGetLastError());
    return GSB_OK;
}

// ---- full-grid statistics through the column kernel -------------------------
// Columns deeper than the register/shared-memory kernels hold (R > 1024) are
// rare (the reference recommends "a few hundreds" of realizations,
// docs/gui.rst:399-404) but must work: every node becomes a K = 1 column of
// K2c (block-wide sort, in global memory when it does not fit on chip), chunk
// by chunk, and the float64 table is transposed into the float32 planes.
namespace {
__global__ void iota_ids_kernel(long long* ids, long long first, long long n) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) ids[i] = first + i;
}
__global__ void table_to_planes_kernel(const double* __restrict__ table, int nplanes,
                                       long long n, float* out, long long out_pitch) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int k = 0; k < nplanes; ++k)
        out[static_cast<long long>(k) * out_pitch + i] = static_cast<float>(table[i * nplanes + k]);
}
}  // namespace

int gsb_launch_k2_deep(gsb_stack* s, const GsbPlanes& pl, float nodata, int64_t n0, int64_t nn,
                       float* d_out, int64_t out_pitch, cudaStream_t st) {
    // up to 32768 values the column kernel sorts in shared memory; deeper columns
    // would make it regrow the launcher scratch this function is using
    GSB_REQUIRE(s->R <= 32768, "full-grid order statistics support R <= 32768 (got %lld)",
                (long long)s->R);
    const int64_t chunk = 32768;
    const size_t ids_bytes = static_cast<size_t>(chunk) * 8;
    void* d = nullptr;
    int rc = gsb_scratch(s, ids_bytes + static_cast<size_t>(chunk) * pl.n * 8, &d);
    if (rc) return rc;
    long long* d_ids = static_cast<long long*>(d);
    double* d_tab = reinterpret_cast<double*>(static_cast<char*>(d) + ids_bytes);
    for (int64_t a = 0; a < nn; a += chunk) {
        const int64_t m = nn - a < chunk ? nn - a : chunk;
        const int blocks = static_cast<int>((m + 255) / 256);
        iota_ids_kernel<<<blocks, 256, 0, st>>>(d_ids, n0 + a, m);
        rc = gsb_launch_k2_columns(s, reinterpret_cast<const int64_t*>(d_ids), static_cast<int>(m),
                                   1, 1, pl, nodata, d_tab, st);
        if (rc) return rc;
        table_to_planes_kernel<<<blocks, 256, 0, st>>>(d_tab, pl.n, m, d_out + a, out_pitch);
        GSB_CUDA(cudaGetLastError());
    }
    return GSB_OK;
}

int gsb_launch_extract(gsb_stack* s, const int64_t* d_node_ids, int64_t ncols, int K,
                       float* d_values, cudaStream_t st) {
    const long long total = ncols * K * s->R;
    if (total <= 0) return GSB_OK;
    long long blocks = (total + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    extract_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(
        s->data, s->pitch, static_cast<int>(s->R), K,
        reinterpret_cast<const long long*>(d_node_ids), ncols, d_values);
    GSB_CUDA(cudaGetLastError());
    return GSB_OK;
}

int gsb_launch_k3_detect(const double* clim, const double* mean, const double* lperc,
                         const double* rperc, const double* fix_a, const double* fix_b,
                         const double* fix_c, int S, int dz, int method, double thr,
                         double nodata, double* filled, double* homog, double* flag,
                         int32_t* detected, cudaStream_t st) {
    if (S <= 0) return GSB_OK;
    // columns a method does not use may be NULL: they are read, never used
    K3Args a{clim, mean, lperc, rperc, fix_a, fix_b ? fix_b : fix_a, fix_c ? fix_c : fix_a, S, dz,
             method, thr, nodata, filled, homog, flag, detected};
    k3_detect_kernel<<<S, 128, 0, st>>>(a);
    GSB_CUDA(cudaGetLastError());
    return GSB_OK;
}

int gsb_launch_k3_detect_table(const double* clim, const double* table, const double* fixtab,
                               const int64_t* zoff, int64_t sstride, int C, int Cfix, int c_mean,
                               int c_lperc, int c_rperc, int c_fa, int c_fb, int c_fc, int S, int dz,
                               int method, double thr, double nodata, double* out,
                               int32_t* detected, cudaStream_t st) {
    if (S <= 0) return GSB_OK;
    K3TabArgs a;
    a.clim = clim; a.table = table; a.fixtab = fixtab;
    a.zoff = reinterpret_cast<const long long*>(zoff); a.sstride = sstride;
    a.C = C; a.Cfix = Cfix;
    a.c_mean = c_mean; a.c_lperc = c_lperc; a.c_rperc = c_rperc;
    a.c_fa = c_fa; a.c_fb = c_fb; a.c_fc = c_fc;
    a.S = S; a.dz = dz; a.method = method; a.thr = thr; a.nodata = nodata;
    a.out = out; a.detected = detected;
    k3_detect_table_kernel<<<S, 128, 0, st>>>(a);
    GSB_CUDA(cudaGetLastError());
    return GSB_OK;
}
