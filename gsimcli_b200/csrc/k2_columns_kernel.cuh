// k2_columns_kernel.cuh -- K2c (station columns) and K3 (detect) kernels; see
// k2_columns.cu for the description and the launchers.  In a header so that
// tests/emul can compile THIS code for the host and run it with one OS thread per
// CUDA thread (tests/emul/k2_columns_threads.cpp; ThreadSanitizer, fuzz against the
// oracle).  Host build: KC_HOST is defined by the includer, which supplies the
// CUDA built-ins (tests/emul/cuda_threads_shim.h; __shared__ becomes static: one
// block runs at a time).
#pragma once
#include "gsb_common.cuh"

#ifndef KC_HOST
#define KC_GLOBAL(NT) __global__ void __launch_bounds__(NT)
#define KC_GLOBAL_PLAIN __global__ void
#define KC_PARAM const __grid_constant__
#define KC_DEV __device__ __forceinline__
#define KC_FN __device__
#else
#define KC_GLOBAL(NT) static void
#define KC_GLOBAL_PLAIN static void
#define KC_PARAM const
#define KC_DEV static inline
#define KC_FN static
#endif

namespace kcol {

constexpr int KC_THREADS = 256;

struct KcArgs {
    const float* data;
    long long pitch;
    int R;
    int K;
    int nplanes;
    int npow2;
    int use_global;
    float nodata;
    const long long* node_ids;   // [ncols][K]
    float* gsort;                // global sort scratch [ncols][npow2] when use_global
    double* table;               // [ncols][nplanes]
    unsigned char kind[GSB_MAX_PLANES];
    double q[GSB_MAX_PLANES];
};

KC_FN double block_sum(double x, double* red) {
    // fixed-order reduction: butterfly inside warps, then warp 0 over the 8 partials
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = x;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < KC_THREADS / 32; ++w) t += red[w];
    return t;
}

KC_GLOBAL(KC_THREADS)
k2_columns_kernel(KC_PARAM KcArgs a) {
#ifndef KC_HOST
    extern __shared__ __align__(16) unsigned char smem_raw[];
#else
    unsigned char* smem_raw = gsb_host_smem();
#endif
    __shared__ double red[KC_THREADS / 32];
    __shared__ int s_n;
    const int col = blockIdx.x;
    float* buf = a.use_global ? (a.gsort + static_cast<long long>(col) * a.npow2)
                              : reinterpret_cast<float*>(smem_raw);
    const long long* ids = a.node_ids + static_cast<long long>(col) * a.K;
    const int total = a.R * a.K;
    const float INF = __int_as_float(0x7f800000);
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();

    // ---- gather (reference order e = i + j*K), nodata/NaN/absent -> +INF
    int myn = 0;
    for (int e = threadIdx.x; e < a.npow2; e += KC_THREADS) {
        float x = INF;
        if (e < total) {
            const int i = e % a.K, j = e / a.K;
            const long long node = ids[i];
            if (node >= 0) {
                const float v = a.data[static_cast<long long>(j) * a.pitch + node];
                if ((v == v) && (v != a.nodata)) { x = v; ++myn; }
            }
        }
        buf[e] = x;
    }
    atomicAdd(&s_n, myn);
    __syncthreads();
    const int n = s_n;
    const double dn = static_cast<double>(n);

    // ---- moments, two-pass in float64 (bn.nanmean / nanvar ddof=1 / pandas G1)
    double s = 0.0;
    for (int e = threadIdx.x; e < total; e += KC_THREADS) {
        const float x = buf[e];
        if (x != INF) s += static_cast<double>(x);
    }
    s = block_sum(s, red);
    const double mean = s / dn;
    double p2 = 0.0, p3 = 0.0, pmax = 0.0;
    for (int e = threadIdx.x; e < total; e += KC_THREADS) {
        const float x = buf[e];
        if (x != INF) {
            const double d = static_cast<double>(x) - mean;
            const double d2 = __dmul_rn(d, d);
            p2 += d2;
            p3 += __dmul_rn(d2, d);
            pmax = fmax(pmax, fabs(static_cast<double>(x)));
        }
    }
    double m2 = block_sum(p2, red);
    double m3 = block_sum(p3, red);
    // max |x| for pandas' fp-error zeroing (nanops.nanskew)
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) pmax = fmax(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = pmax;
    __syncthreads();
    double max_abs = 0.0;
#pragma unroll
    for (int w = 0; w < KC_THREADS / 32; ++w) max_abs = fmax(max_abs, red[w]);
    __syncthreads();

    // ---- block-wide bitonic sort (ascending; +INF padding sorts last)
    for (int k = 2; k <= a.npow2; k <<= 1) {
        for (int j = k >> 1; j >= 1; j >>= 1) {
            for (int t = threadIdx.x; t < (a.npow2 >> 1); t += KC_THREADS) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const bool asc = (i & k) == 0;
                const float x = buf[i], y = buf[l];
                const bool sw = asc ? (x > y) : (x < y);
                if (sw) { buf[i] = y; buf[l] = x; }
            }
            __syncthreads();
        }
    }

    // ---- planes
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    const double var = (n > 1) ? m2 / (dn - 1.0) : qnan;
    const double sd = sqrt(var);
    const double eps = 2.220446049250313e-16;
    const double t1 = eps * max_abs;
    double m2z = (fabs(m2) < t1 * t1 * dn) ? 0.0 : m2;
    double m3z = (fabs(m3) < t1 * t1 * t1 * dn) ? 0.0 : m3;
    for (int kx = threadIdx.x; kx < a.nplanes; kx += KC_THREADS) {
        double r = qnan;
        if (n > 0) {
            switch (a.kind[kx]) {
            case GSB_PK_MEAN: r = mean; break;
            case GSB_PK_VAR: r = var; break;
            case GSB_PK_STD: r = sd; break;
            case GSB_PK_COEFVAR: r = sd / mean * 100.0; break;
            case GSB_PK_SKEW:
                if (n < 3) r = qnan;
                else if (m2z == 0.0) r = 0.0;
                else r = (dn * sqrt(dn - 1.0) / (dn - 2.0)) * (m3z / pow(m2z, 1.5));
                break;
            case GSB_PK_QUANT: {
                const double h = __dmul_rn(dn - 1.0, a.q[kx]);
                if (h >= dn - 1.0) r = static_cast<double>(buf[n - 1]);
                else if (h < 0.0) r = static_cast<double>(buf[0]);
                else {
                    const double fl = floor(h);
                    const int lo = static_cast<int>(fl);
                    const double t = h - fl;
                    const double x0 = static_cast<double>(buf[lo]);
                    const double x1 = static_cast<double>(buf[lo + 1]);
                    const double d = x1 - x0;
                    r = (t >= 0.5) ? __dsub_rn(x1, __dmul_rn(d, 1.0 - t))
                                   : __dadd_rn(x0, __dmul_rn(d, t));
                }
            } break;
            case GSB_PK_MODE: {
                // longest run of equal values, ties -> smallest (serial; S*dz columns only)
                float best = buf[0];
                int bestc = 0, run = 0;
                float prev = buf[0];
                for (int e = 0; e < n; ++e) {
                    const float x = buf[e];
                    run = (x == prev) ? run + 1 : 1;
                    prev = x;
                    if (run > bestc) { bestc = run; best = x; }
                }
                r = static_cast<double>(best);
            } break;
            default: break;
            }
        }
        a.table[static_cast<long long>(col) * a.nplanes + kx] = r;
    }
}

// raw extraction: values[col][i][j] for neighbour i, realization j
KC_GLOBAL_PLAIN
extract_kernel(const float* data, long long pitch, int R, int K,
                               const long long* node_ids, long long ncols, float* values) {
    const long long total = ncols * K * R;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int j = static_cast<int>(e % R);
        const long long ci = e / R;
        const long long node = node_ids[ci];
        values[e] = node >= 0 ? data[static_cast<long long>(j) * pitch + node]
                              : __int_as_float(0x7fc00000);
    }
}

struct K3Args {
    const double *clim, *mean, *lperc, *rperc, *fix_a, *fix_b, *fix_c;
    int S, dz, method;
    double thr, nodata;
    double *filled, *homog, *flag;
    int32_t* detected;
};

// homog.py:191-239 for one (station, time step): fill, test, substitute
struct K3Out { double filled, homog, flag; int hw; };

KC_DEV K3Out k3_apply(double c, double mean, double lo, double hi, double fa,
                                          double fb, double fc, int method, double thr,
                                          double nodata) {
    // homog.py:200-201 -- NaN observations take the local mean
    if (c != c) c = mean;
    // homog.py:204 -- ~between(lperc, rperc), inclusive; NaN compares False.
    // The bounds come from float32 realization values, the observation from
    // float64 text: an observation that TIES a bound in the reference (both
    // parsed from the same decimal, e.g. a DSS clamp value 0.7) must tie here
    // too, so the observation is compared at the stack's precision
    // (f32(0.7) == f32(0.7), whereas 0.7 > f32(0.7) would flag it).
    const double cq = static_cast<double>(static_cast<float>(c));
    const bool inside = (cq >= lo) && (cq <= hi);
    const bool hw = !inside;
    double fix;
    switch (method) {
    case GSB_METHOD_SKEWNESS:      // np.where(skew > thr, median, mean)
        fix = (fc > thr) ? fa : fb;
        break;
    case GSB_METHOD_PERCENTILE:    // np.where(clim > rperc, rperc', lperc')
        fix = (cq > hi) ? fa : fb;
        break;
    default:
        fix = fa;
    }
    K3Out o;
    o.filled = c;
    o.homog = hw ? fix : c;        // clim.where(~hom_where, fixvalues)
    o.flag = hw ? c : nodata;      // clim.where(hom_where, nodata)
    o.hw = hw ? 1 : 0;
    return o;
}

// one CTA per station; thread per time step
KC_GLOBAL(128) k3_detect_kernel(KC_PARAM K3Args a) {
    __shared__ int s_det;
    const int s = blockIdx.x;
    if (threadIdx.x == 0) s_det = 0;
    __syncthreads();
    int det = 0;
    for (int t = threadIdx.x; t < a.dz; t += blockDim.x) {
        const long long e = static_cast<long long>(s) * a.dz + t;
        const K3Out o = k3_apply(a.clim[e], a.mean[e], a.lperc[e], a.rperc[e], a.fix_a[e],
                                 a.fix_b[e], a.fix_c[e], a.method, a.thr, a.nodata);
        a.filled[e] = o.filled;
        a.homog[e] = o.homog;
        a.flag[e] = o.flag;
        det += o.hw;
    }
    atomicAdd(&s_det, det);
    __syncthreads();
    if (threadIdx.x == 0) a.detected[s] = s_det;
}

// the same step reading the K2c station tables where they lie on the device:
// row (s, z) of a table is at ((zoff[z] + s * sstride) * C); one table for the
// detection interval + mean, one for the substitute (they may be the same).
// zoff / sstride describe either a plain [S][dz][C] table (zoff[z] = z,
// sstride = dz) or the per-rank slabs an all-gather leaves behind
// ([world][S][width][C]: zoff[z] = rank(z) * S * width + z - z0(rank), sstride = width).
struct K3TabArgs {
    const double* clim;
    const double* table;
    const double* fixtab;
    const long long* zoff;
    long long sstride;
    int C, Cfix;
    int c_mean, c_lperc, c_rperc, c_fa, c_fb, c_fc;
    int S, dz, method;
    double thr, nodata;
    double* out;            // [3][S][dz]: filled, homog, flag
    int32_t* detected;
};

KC_GLOBAL(128) k3_detect_table_kernel(KC_PARAM K3TabArgs a) {
    __shared__ int s_det;
    const int s = blockIdx.x;
    if (threadIdx.x == 0) s_det = 0;
    __syncthreads();
    int det = 0;
    const long long plane = static_cast<long long>(a.S) * a.dz;
    for (int t = threadIdx.x; t < a.dz; t += blockDim.x) {
        const long long e = static_cast<long long>(s) * a.dz + t;
        const long long row = a.zoff[t] + static_cast<long long>(s) * a.sstride;
        const double* tab = a.table + row * a.C;
        const double* fix = a.fixtab + row * a.Cfix;
        const K3Out o = k3_apply(a.clim[e], tab[a.c_mean], tab[a.c_lperc], tab[a.c_rperc],
                                 fix[a.c_fa], fix[a.c_fb], a.c_fc >= 0 ? tab[a.c_fc] : 0.0,
                                 a.method, a.thr, a.nodata);
        a.out[e] = o.filled;
        a.out[plane + e] = o.homog;
        a.out[2 * plane + e] = o.flag;
        det += o.hw;
    }
    atomicAdd(&s_det, det);
    __syncthreads();
    if (threadIdx.x == 0) a.detected[s] = s_det;
}

}  // namespace kcol
