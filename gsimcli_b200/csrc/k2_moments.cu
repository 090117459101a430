// k2_moments.cu -- K2m: streaming moments when no order statistic is requested
//
// GridFiles.stats(lmean/lvar/lstd/lcoefvar/lskew) (GSIMCLI/tools/grid.py:680-698)
// without median/percentiles: a pure HBM-bound stream.  One thread owns one
// node and walks the R rows; a warp reads 128 contiguous bytes of every row
// (the stack is realization-major, so no transpose and no shared memory).
// Shifted one-pass sums around a pivot (the valid value nearest to the mean of
// the first 16 rows, which stay in registers so nothing is re-read): differences and products in float32,
// 16-row partial sums flushed into float64 accumulators.  Algorithmic traffic:
// 4 B per node-realization + 4 B per node and output map.
#include "gsb_common.cuh"

namespace {

constexpr int KM_THREADS = 256;
constexpr int KM_HEAD = 16;

struct KmArgs {
    const float* data;
    long long pitch;
    int R;
    int nplanes;
    float nodata;
    long long n0;
    long long nn;
    float* out;
    long long out_pitch;
    unsigned char kind[GSB_MAX_PLANES];
};

__device__ __forceinline__ float ldg_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(KM_THREADS)
k2_moments_kernel(const __grid_constant__ KmArgs a) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long node = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
         node < a.nn; node += stride) {
        const float* col = a.data + a.n0 + node;
        // ---- head rows: pivot
        float head[KM_HEAD];
        float hsum = 0.f;
        int hc = 0;
        const int nh = a.R < KM_HEAD ? a.R : KM_HEAD;
#pragma unroll
        for (int r = 0; r < KM_HEAD; ++r) {
            head[r] = (r < nh) ? ldg_stream(col + r * a.pitch) : a.nodata;
            const bool ok = (head[r] == head[r]) && (head[r] != a.nodata);
            hsum += ok ? head[r] : 0.f;
            hc += ok ? 1 : 0;
        }
        // pivot: the valid head value nearest to the head mean -- an element (a
        // constant column gives d == 0 exactly), never an outlier, and within a
        // fraction of sigma of the column mean (the float32 shifted sums lose
        // |pivot - mean| / sigma digits in the third moment)
        float piv = 0.f;
        {
            const float hm = hsum / static_cast<float>(hc);
            float bestd = __int_as_float(0x7f800000);
            bool found = false;
#pragma unroll
            for (int r = 0; r < KM_HEAD; ++r) {
                const bool ok = (head[r] == head[r]) && (head[r] != a.nodata);
                const float d = fabsf(head[r] - hm);
                const bool take = ok && (!found || d < bestd);     // NaN distances: first valid one
                piv = take ? head[r] : piv;
                bestd = take ? d : bestd;
                found = found || ok;
            }
        }
        double S1 = 0.0, S2 = 0.0, S3 = 0.0;
        int n = 0;
        float vmin = __int_as_float(0x7f800000), vmax = __int_as_float(0xff800000);
        {
            float s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int r = 0; r < KM_HEAD; ++r) {
                const bool ok = (head[r] == head[r]) && (head[r] != a.nodata);
                const float d = ok ? head[r] - piv : 0.f;
                const float d2 = d * d;
                s1 += d; s2 += d2; s3 = fmaf(d2, d, s3);
                vmin = ok ? fminf(vmin, head[r]) : vmin;
                vmax = ok ? fmaxf(vmax, head[r]) : vmax;
            }
            S1 = s1; S2 = s2; S3 = s3; n = hc;
        }
        // ---- body, 16 rows per flush
        int r = KM_HEAD;
        for (; r + 16 <= a.R; r += 16) {
            float x[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) x[k] = ldg_stream(col + (r + k) * a.pitch);
            if (hc == 0) {            // no pivot yet: take the first valid value
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    if (hc == 0 && (x[k] == x[k]) && (x[k] != a.nodata)) { piv = x[k]; hc = 1; }
            }
            float s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const bool ok = (x[k] == x[k]) && (x[k] != a.nodata);
                const float d = ok ? x[k] - piv : 0.f;
                const float d2 = d * d;
                s1 += d; s2 += d2; s3 = fmaf(d2, d, s3);
                n += ok ? 1 : 0;
                vmin = ok ? fminf(vmin, x[k]) : vmin;
                vmax = ok ? fmaxf(vmax, x[k]) : vmax;
            }
            S1 += s1; S2 += s2; S3 += s3;
        }
        {
            float s1 = 0.f, s2 = 0.f, s3 = 0.f;
            for (; r < a.R; ++r) {
                const float x = ldg_stream(col + r * a.pitch);
                const bool ok = (x == x) && (x != a.nodata);
                if (ok && hc == 0) { piv = x; hc = 1; }
                const float d = ok ? x - piv : 0.f;
                const float d2 = d * d;
                s1 += d; s2 += d2; s3 = fmaf(d2, d, s3);
                n += ok ? 1 : 0;
                vmin = ok ? fminf(vmin, x) : vmin;
                vmax = ok ? fmaxf(vmax, x) : vmax;
            }
            S1 += s1; S2 += s2; S3 += s3;
        }
        // ---- finish in float64
        const double qnan = __longlong_as_double(0x7ff8000000000000LL);
        const double dn = static_cast<double>(n);
        const double delta = S1 / dn;
        const double mean = static_cast<double>(piv) + delta;
        const bool constant = (vmin == vmax);
        double m2 = S2 - S1 * delta;
        double m3 = S3 - 3.0 * delta * S2 + 2.0 * dn * delta * delta * delta;
        if (constant) { m2 = 0.0; m3 = 0.0; }
        if (m2 < 0.0) m2 = 0.0;
        const double var = (n > 1) ? m2 / (dn - 1.0) : qnan;
        const double sd = sqrt(var);
        for (int k = 0; k < a.nplanes; ++k) {
            double v = qnan;
            if (n > 0) {
                switch (a.kind[k]) {
                case GSB_PK_MEAN: v = mean; break;
                case GSB_PK_VAR: v = var; break;
                case GSB_PK_STD: v = sd; break;
                case GSB_PK_COEFVAR: v = sd / mean * 100.0; break;
                case GSB_PK_SKEW:
                    v = (n < 3) ? qnan : (m2 == 0.0) ? 0.0
                        : (dn * sqrt(dn - 1.0) / (dn - 2.0)) * (m3 / (m2 * sqrt(m2)));
                    break;
                default: break;
                }
            }
            a.out[k * a.out_pitch + node] = static_cast<float>(v);
        }
    }
}

}  // namespace

int gsb_launch_k2_moments(gsb_stack* s, const GsbPlanes& pl, float nodata, int64_t n0,
                          int64_t nn, float* d_out, int64_t out_pitch, cudaStream_t st) {
    if (nn <= 0) return GSB_OK;
    KmArgs a;
    memset(&a, 0, sizeof(a));
    a.data = s->data;
    a.pitch = s->pitch;
    a.R = static_cast<int>(s->R);
    a.nplanes = pl.n;
    a.nodata = nodata;
    a.n0 = n0;
    a.nn = nn;
    a.out = d_out;
    a.out_pitch = out_pitch;
    for (int i = 0; i < pl.n; ++i) a.kind[i] = pl.kind[i];
    long long blocks = (nn + KM_THREADS - 1) / KM_THREADS;
    const long long cap = static_cast<long long>(s->sm_count) * 8 * 4;
    if (blocks > cap) blocks = cap;
    k2_moments_kernel<<<static_cast<int>(blocks), KM_THREADS, 0, st>>>(a);
    GSB_CUDA(cudaGetLastError());
    return GSB_OK;
}
