// k2_csort_threads.cpp -- TEST INFRASTRUCTURE: the K2 counting-sort kernel (the
// dominant kernel of the path; gsimcli_b200/csrc/k2_csort_kernel.cuh, the very
// function nvcc compiles) built for the host and run with one OS thread per CUDA
// thread on the model of cuda_threads_shim.h.  Stand-alone driver: planes must
// equal the serial emulation of the register-tile kernel (k2_rsel_emul.cpp, itself
// pinned to the oracle by tests/test_k2_rsel_emul.py) -- bit for bit for the order
// statistics and the mode, 1e-5 for the moments -- in both TMA emulation modes, on
// sub-ranges that are not tile-aligned, with more tiles than CTAs, with a tile list.
// Under -fsanitize=thread this is the closest substitute for compute-sanitizer's
// racecheck on the kernel's ticket / mbarrier / reader-count hand-off that the GPU
// pool allows (it refuses the real tool: profiles/r02_sanitizer_memcheck.log).
// Never linked into libgsimcli_b200.so.
//   g++ -O1 -g -std=c++17 -pthread -ffp-contract=off -I/usr/local/cuda/include [-fsanitize=thread] ...
#include <stdio.h>
#include "../../gsimcli_b200/csrc/gsb_common.cuh"
#include "cuda_threads_shim.h"
#define CS_HOST 1
#include "../../gsimcli_b200/csrc/k2_csort_kernel.cuh"
#include "k2_rsel_emul.cpp"

using namespace csort_k;

template <int IPT, bool MODE>
static void run_grid(const CUtensorMap& tmap, const K2Args& a, int grid, int tma_mode) {
    constexpr int HALO = 10;
    constexpr int CS_WARPS = CsWarps<IPT>::v;
    constexpr int NPAD = IPT * 32;
    // the launcher's formula (k2_grid_csort.cu, cs_launch_one)
    const size_t smem = static_cast<size_t>(NPAD) * 128
        + static_cast<size_t>(CS_WARPS) * K2_CS_REGION_WORDS(IPT, HALO) * 4 + 16
        + GSB_MAX_PLANES * 12 + CS_WARPS * 100 * 4;
    for (int b = 0; b < grid; ++b)
        shim_run_block(b, grid, CS_WARPS * 32, smem, tma_mode,
                       [&]() { k2_csort_kernel<IPT, MODE, HALO>(tmap, a); });
}

static int csort_threads(const float* data, long long pitch, long long N, int R, float nodata, long long n0,
                         long long nn, int nplanes, const unsigned char* kind, const double* q, float* out,
                         long long out_pitch, int grid, int tma_mode, const int* tile_list) {
    K2Args args;
    memset(&args, 0, sizeof(args));
    args.R = R;
    args.nchunks = (R + 31) / 32;
    args.nplanes = nplanes;
    args.nodata = nodata;
    const long long head = n0 & 3;
    args.n0 = n0 - head;
    args.nn = nn + head;
    args.skip = head;
    args.ntiles = (args.nn + 31) / 32;
    args.out = out - head;
    args.out_pitch = out_pitch;
    args.tile_list = tile_list;
    bool mode = false;
    for (int i = 0; i < nplanes; ++i) {
        args.kind[i] = kind[i];
        args.q[i] = q[i];
        mode = mode || kind[i] == GSB_PK_MODE;
    }
    int ipt = 1;
    while (ipt * 32 < R) ipt <<= 1;
    ShimTmap d;
    d.data = data;
    d.pitch = pitch;
    d.N = N;
    d.R = R;
    d.box_rows = ipt * 32 < 256 ? ipt * 32 : 256;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    memcpy(&tmap, &d, sizeof(d));
    if (grid > args.ntiles) grid = static_cast<int>(args.ntiles);
#define GO(I) do { if (mode) run_grid<I, true>(tmap, args, grid, tma_mode); else run_grid<I, false>(tmap, args, grid, tma_mode); } while (0)
    switch (ipt) {
    case 2: GO(2); break;
    case 4: GO(4); break;
    case 8: GO(8); break;
    case 16: GO(16); break;
    case 32: GO(32); break;
    default: return -1;
    }
    return 0;
}

// library entry (tests/test_k2_csort_emul.py): same arguments as csort_threads
extern "C" int k2_csort_emulate_threads(const float* data, long long pitch, long long N, int R, float nodata,
                                        long long n0, long long nn, int nplanes, const unsigned char* kind,
                                        const double* q, float* out, long long out_pitch, int grid,
                                        int tma_mode) {
    if (R <= 32 || R > 1024 || pitch < 32 || (pitch & 31)) return -1;
    return csort_threads(data, pitch, N, R, nodata, n0, nn, nplanes, kind, q, out, out_pitch, grid, tma_mode,
                         nullptr);
}

#ifndef CSORT_LIB
static unsigned long long rng_state = 0x2545F4914F6CDD1Dull;
static unsigned rnd() {
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return static_cast<unsigned>(rng_state >> 32);
}

static int one(int R, int grid, long long n0, long long nn, bool with_mode, bool use_list) {
    const long long N = 150, pitch = 160;
    const float nodata = -999.9f;
    std::vector<float> data(static_cast<size_t>(R) * pitch, 1.0e30f);
    for (int r = 0; r < R; ++r)
        for (long long c = 0; c < N; ++c) {
            const unsigned u = rnd();
            float v = 400.0f + static_cast<float>(u % 8000) / 16.0f + static_cast<float>(c % 7) * 20.0f;
            if (v > 800.0f) v = 800.0f;                                   // clamp atom (maximum)
            if (c % 5 == 0 && v < 450.0f) v = 450.0f;                     // ... and minimum
            if ((u >> 20) % 97 == 0) v = nodata;
            data[static_cast<size_t>(r) * pitch + c] = v;
        }
    for (int r = 0; r < R; ++r) {
        data[static_cast<size_t>(r) * pitch + 11] = 512.25f;              // constant column
        data[static_cast<size_t>(r) * pitch + 12] = nodata;               // empty column
        data[static_cast<size_t>(r) * pitch + 13] = static_cast<float>(rnd() % 4) * 0.5f;   // heavy ties: big bins
        if (r < R / 3 || r < 33) data[static_cast<size_t>(r) * pitch + 40] = nodata;        // first rows missing
        // ... and a narrow column far from zero (mean / sigma = 70): what a pivot of 0 costs
        data[static_cast<size_t>(r) * pitch + 41] = r < 40 ? nodata : 800.0f + static_cast<float>(rnd() % 641) / 16.0f - 20.0f;
    }
    data[static_cast<size_t>(1) * pitch + 20] = 3.0e8f;                   // outlier: the bin map collapses
    std::vector<unsigned char> kind = {0, 1, 2, 3, 4, 5, 1, 1, 1, 1, 1};
    std::vector<double> q = {0, .5, 0, 0, 0, 0, .025, .975, .05, .5, .95};
    if (with_mode) { kind.push_back(6); q.push_back(0); }
    const int nplanes = static_cast<int>(kind.size());
    std::vector<float> ref(static_cast<size_t>(nplanes) * N, -1.f);
    std::vector<int> fl(64, 0);
    const int nf = k2_rsel_emulate(data.data(), pitch, N, R, nodata, nplanes, kind.data(), q.data(), ref.data(), N,
                                   32, fl.data());
    if (nf < 0) { printf("reference emulation refused R=%d\n", R); return 1; }
    std::vector<char> skipcol(N, 0);                                      // tiles the register tile hands over
    for (int k = 0; k < nf; ++k)
        for (int c = 0; c < 32 && fl[k] * 32 + c < N; ++c) skipcol[fl[k] * 32 + c] = 1;
    int bad = 0;
    for (int tma_mode = 0; tma_mode < 2; ++tma_mode) {
        std::vector<float> out(static_cast<size_t>(nplanes) * nn, -2.f);
        std::vector<int> list;
        const long long ntiles = (nn + (n0 & 3) + 31) / 32;
        if (use_list) {                                                    // every other tile, out of order
            list.push_back(0);
            for (long long t = ntiles - 1; t >= 0; t -= 2) { list.push_back(static_cast<int>(t)); ++list[0]; }
        }
        if (csort_threads(data.data(), pitch, N, R, nodata, n0, nn, nplanes, kind.data(), q.data(), out.data(), nn,
                          grid, tma_mode, use_list ? list.data() : nullptr) != 0) { printf("unsupported\n"); return 1; }
        for (long long c = 0; c < nn; ++c) {
            const long long tile = (c + (n0 & 3)) / 32;
            bool listed = !use_list;
            for (size_t k = 1; k < list.size(); ++k) listed = listed || list[k] == tile;
            for (int pl = 0; pl < nplanes; ++pl) {
                const float g = out[static_cast<size_t>(pl) * nn + c];
                if (!listed) { if (g != -2.f) { if (bad < 5) printf("unlisted tile written\n"); ++bad; } continue; }
                if (skipcol[n0 + c]) continue;
                const float e = ref[static_cast<size_t>(pl) * N + n0 + c];
                const bool exact = kind[pl] == 1 || kind[pl] == 6;
                bool ok;
                if (g != g || e != e) ok = (g != g) && (e != e);
                else if (exact) ok = memcmp(&g, &e, 4) == 0;
                else ok = fabsf(g - e) <= 1e-5f * fabsf(e) + (kind[pl] == 2 ? 1e-5f : 1e-6f);
                if (!ok) {
                    if (bad < 8) printf("R=%d mode=%d plane %d (kind %d) col %lld: %.9g vs %.9g\n", R, tma_mode, pl,
                                        kind[pl], n0 + c, g, e);
                    ++bad;
                }
            }
        }
    }
    return bad;
}

int main(int argc, char** argv) {
    const bool quick = argc > 1;              // under ThreadSanitizer: fewer shapes
    int bad = 0;
    bad += one(1000, 2, 6, 140, true, false);     // IPT 32, 16 warps: 5 tiles on 2 CTAs, head = 2, ragged end
    bad += one(200, 2, 0, 150, true, false);      // IPT 8, 32 warps
    if (!quick) {
        bad += one(1000, 3, 0, 150, false, false);
        bad += one(500, 2, 34, 100, true, false); // IPT 16, 24 warps
        bad += one(100, 2, 0, 150, true, false);  // IPT 4
        bad += one(1000, 2, 0, 150, true, true);  // tile list (the register-tile kernel's hand-over)
    }
    if (bad) { printf("%d mismatches\n", bad); return 1; }
    printf("clean\n");
    return 0;
}
#endif  // CSORT_LIB
