// k2_columns_threads.cpp -- TEST INFRASTRUCTURE: K2c (station columns: gather of the
// disc values, float64 moments, block-wide bitonic sort, NumPy _lerp) and K3
// (detect / correct) -- gsimcli_b200/csrc/k2_columns_kernel.cuh, the kernels nvcc
// compiles -- built for the host and run with one OS thread per CUDA thread
// (tests/emul/cuda_threads_shim.h).  Library entry points for the Python fuzz
// against the oracle (tests/test_k2_columns_emul.py); under -fsanitize=thread the
// barriers of the block-wide sort and of the fixed-order reductions are checked.
// Never linked into libgsimcli_b200.so.
#include "../../gsimcli_b200/csrc/gsb_common.cuh"
#include "cuda_threads_shim.h"
#undef __shared__
#define __shared__ static
#define KC_HOST 1
#include "../../gsimcli_b200/csrc/k2_columns_kernel.cuh"

using namespace kcol;

// node_ids: [ncols][K] (-1 = absent), table: [ncols][nplanes] float64.  Shared-memory
// sort for npow2 * 4 <= 200 KB, else the global scratch -- the launcher's rule; force_global
// exercises the other path on small inputs.
extern "C" int k2c_emulate(const float* data, long long pitch, int R, int K, const long long* node_ids,
                           long long ncols, float nodata, int nplanes, const unsigned char* kind,
                           const double* q, double* table, int force_global) {
    const long long total = static_cast<long long>(R) * K;
    if (total > (1 << 22) || ncols <= 0) return -1;
    int npow2 = 32;
    while (npow2 < total) npow2 <<= 1;
    KcArgs a;
    memset(&a, 0, sizeof(a));
    a.data = data;
    a.pitch = pitch;
    a.R = R;
    a.K = K;
    a.nplanes = nplanes;
    a.npow2 = npow2;
    a.nodata = nodata;
    a.node_ids = node_ids;
    a.table = table;
    for (int i = 0; i < nplanes; ++i) { a.kind[i] = kind[i]; a.q[i] = q[i]; }
    size_t smem = static_cast<size_t>(npow2) * 4;
    a.use_global = (smem > 200 * 1024) || force_global;
    std::vector<float> scratch;
    if (a.use_global) {
        scratch.assign(static_cast<size_t>(ncols) * npow2, -1.f);
        a.gsort = scratch.data();
        smem = 16;
    }
    for (long long b = 0; b < ncols; ++b)
        shim_run_block(static_cast<int>(b), static_cast<int>(ncols), KC_THREADS, smem, 0,
                       [&]() { k2_columns_kernel(a); });
    return 0;
}

extern "C" int k3_emulate(const double* clim, const double* mean, const double* lperc, const double* rperc,
                          const double* fa, const double* fb, const double* fc, int S, int dz, int method,
                          double thr, double nodata, double* filled, double* homog, double* flag,
                          int32_t* detected) {
    K3Args a;
    memset(&a, 0, sizeof(a));
    a.clim = clim; a.mean = mean; a.lperc = lperc; a.rperc = rperc;
    a.fix_a = fa; a.fix_b = fb; a.fix_c = fc;
    a.S = S; a.dz = dz; a.method = method; a.thr = thr; a.nodata = nodata;
    a.filled = filled; a.homog = homog; a.flag = flag; a.detected = detected;
    for (int s = 0; s < S; ++s)
        shim_run_block(s, S, 128, 16, 0, [&]() { k3_detect_kernel(a); });
    return 0;
}

#ifdef KC_MAIN
// stand-alone run for ThreadSanitizer: a few columns through both sort paths
#include <stdio.h>
int main() {
    const int R = 60, K = 5, ncols = 6, pitch = 64;
    std::vector<float> data(static_cast<size_t>(R) * pitch);
    unsigned long long st = 88172645463325252ull;
    for (auto& v : data) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; v = static_cast<float>((st >> 40) % 4000) / 16.0f; }
    for (int r = 0; r < R; r += 7) data[static_cast<size_t>(r) * pitch + 3] = -999.9f;
    std::vector<long long> ids(ncols * K);
    for (int c = 0; c < ncols; ++c)
        for (int i = 0; i < K; ++i) ids[c * K + i] = (i == 4 && c % 2) ? -1 : (c * 7 + i * 3) % 60;
    const unsigned char kind[9] = {0, 1, 2, 3, 4, 5, 1, 1, 6};
    const double q[9] = {0, .5, 0, 0, 0, 0, .025, .975, 0};
    std::vector<double> t0(ncols * 9), t1(ncols * 9);
    k2c_emulate(data.data(), pitch, R, K, ids.data(), ncols, -999.9f, 9, kind, q, t0.data(), 0);
    k2c_emulate(data.data(), pitch, R, K, ids.data(), ncols, -999.9f, 9, kind, q, t1.data(), 1);
    int bad = 0;
    for (size_t i = 0; i < t0.size(); ++i)
        if (memcmp(&t0[i], &t1[i], 8) != 0 && !(t0[i] != t0[i] && t1[i] != t1[i])) ++bad;
    const int S = 3, dz = 40;
    std::vector<double> clim(S * dz), mean(S * dz, 10.0), lo(S * dz, 5.0), hi(S * dz, 15.0), fa(S * dz, 11.0),
        filled(S * dz), homog(S * dz), flag(S * dz);
    for (int i = 0; i < S * dz; ++i) clim[i] = (i % 9 == 0) ? NAN : 4.0 + (i % 13);
    std::vector<int32_t> det(S);
    k3_emulate(clim.data(), mean.data(), lo.data(), hi.data(), fa.data(), fa.data(), fa.data(), S, dz, 1, 0.0, -999.9,
               filled.data(), homog.data(), flag.data(), det.data());
    int cnt = 0;
    for (int i = 0; i < S * dz; ++i) cnt += (flag[i] != -999.9) ? 1 : 0;
    if (cnt != det[0] + det[1] + det[2]) ++bad;
    printf(bad ? "%d mismatches\n" : "clean\n", bad);
    return bad ? 1 : 0;
}
#endif
