// k2_rsel_threads.cpp -- TEST INFRASTRUCTURE: the tile loop of the K2 register-tile
// kernel (gsimcli_b200/csrc/k2_rsel_kernel.cuh: the kernel function nvcc compiles,
// the default for the reference's own call shape) built for the host and run with
// one OS thread per CUDA thread on the model of cuda_threads_shim.h.  Planes must
// equal the serial emulation (k2_rsel_emul.cpp) bit for bit, flagged tiles must be
// the same; under -fsanitize=thread a missing barrier or a region reused too early
// (the sort buffer aliases the partial sums, the registers take the next tile while
// the current one is finished) is a data race.  Never linked into libgsimcli_b200.so.
#include <stdio.h>
#include <algorithm>
#include "../../gsimcli_b200/csrc/gsb_common.cuh"
#include "cuda_threads_shim.h"
#define RSK_HOST 1
#include "../../gsimcli_b200/csrc/k2_rsel_kernel.cuh"
#include "k2_rsel_emul.cpp"

using namespace rsel;

template <int NW, int IPT, bool MODE>
static void run_grid(RselLaunch L, int grid) {
    // the launcher's formula (k2_grid_rsel.cu, rsel_smem)
    const int need_rows = L.a.R * COLS, need_alias = 6 * NW * COLS;
    const int words = ((need_rows > need_alias ? need_rows : need_alias) + 3) & ~3;
    L.sort_words = words;
    const size_t fixed = (static_cast<size_t>(NCELLS) * COLS * 4 + 6 * (NW / 8) * COLS * 8 + COLS * 8 +
                          GSB_MAX_PLANES * COLS * 4 + NW * COLS * 4 + CI_WORDS * COLS * 4 + 15) & ~size_t(15);
    const size_t smem = fixed + static_cast<size_t>(words) * 4;
    for (int b = 0; b < grid; ++b)
        shim_run_block(b, grid, NW * 32, smem, 0, [&]() { k2_rsel_kernel<NW, IPT, MODE>(L); });
}

template <int NW, bool MODE>
static int run_nw(const RselLaunch& L, int grid) {
    const int ipt = (L.a.R + NW - 1) / NW;
    if (NW == 32 && ipt <= 2) { run_grid<NW, 2, MODE>(L, grid); return 0; }
    if (ipt <= 4) { run_grid<NW, 4, MODE>(L, grid); return 0; }
    if (ipt <= 8) { run_grid<NW, 8, MODE>(L, grid); return 0; }
    if (ipt <= 16) { run_grid<NW, 16, MODE>(L, grid); return 0; }
    if (ipt <= 32) { run_grid<NW, 32, MODE>(L, grid); return 0; }
    if (NW <= 16 && ipt <= 64) { run_grid<NW <= 16 ? NW : 16, 64, MODE>(L, grid); return 0; }
    return -1;
}

static unsigned long long rng_state = 0xD1B54A32D192ED03ull;
static unsigned rnd() {
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return static_cast<unsigned>(rng_state >> 32);
}

static int one(int R, int nw, int grid, long long n0, long long nn, bool with_mode) {
    const long long N = 150, pitch = 160;
    const float nodata = -999.9f;
    std::vector<float> data(static_cast<size_t>(R) * pitch, 1.0e30f);
    for (int r = 0; r < R; ++r)
        for (long long c = 0; c < N; ++c) {
            const unsigned u = rnd();
            float v = 400.0f + static_cast<float>(u % 8000) / 16.0f + static_cast<float>(c % 7) * 20.0f;
            if (v > 800.0f) v = 800.0f;
            if ((u >> 20) % 97 == 0) v = nodata;
            data[static_cast<size_t>(r) * pitch + c] = v;
        }
    for (int r = 0; r < R; ++r) {
        data[static_cast<size_t>(r) * pitch + 11] = 512.25f;
        data[static_cast<size_t>(r) * pitch + 12] = nodata;
        data[static_cast<size_t>(r) * pitch + 77] = static_cast<float>(rnd() % 4) * 0.5f;   // heavy ties: flagged tile
        if (r < R / 3) data[static_cast<size_t>(r) * pitch + 40] = nodata;
    }
    std::vector<unsigned char> kind = {0, 1, 2, 3, 4, 5, 1, 1};
    std::vector<double> q = {0, .5, 0, 0, 0, 0, .025, .975};
    if (with_mode) { kind.push_back(1); q.push_back(0.3); kind.push_back(6); q.push_back(0); }
    const int nplanes = static_cast<int>(kind.size());
    std::vector<float> ref(static_cast<size_t>(nplanes) * N, -1.f);
    std::vector<int> fl(64, 0);
    const int nf = k2_rsel_emulate(data.data(), pitch, N, R, nodata, nplanes, kind.data(), q.data(), ref.data(), N,
                                   nw, fl.data());
    if (nf < 0) return 1;
    RselLaunch L;
    memset(&L, 0, sizeof(L));
    L.a.R = R;
    L.a.nplanes = nplanes;
    L.a.nodata = nodata;
    int nq = 0;
    for (int i = 0; i < nplanes; ++i) {
        L.a.kind[i] = kind[i];
        L.a.q[i] = q[i];
        if (kind[i] == PK_QUANT) L.a.qplane[nq++] = static_cast<unsigned char>(i);
    }
    L.a.nq = nq;
    std::vector<float> out(static_cast<size_t>(nplanes) * nn, -2.f);
    std::vector<int> flagged(64, 0);
    const long long head = n0 & 3;
    L.data = data.data();
    L.pitch = pitch;
    L.N = N;
    L.n0 = n0 - head;
    L.nn = nn + head;
    L.skip = head;
    L.ntiles = (L.nn + 31) / 32;
    L.out = out.data() - head;
    L.out_pitch = nn;
    L.flagged = flagged.data();
    if (grid > L.ntiles) grid = static_cast<int>(L.ntiles);
    int rc;
    if (nw == 16) rc = with_mode ? run_nw<16, true>(L, grid) : run_nw<16, false>(L, grid);
    else rc = with_mode ? run_nw<32, true>(L, grid) : run_nw<32, false>(L, grid);
    if (rc) return 1;
    // columns of tiles the threaded run flagged are not compared (the counting sort redoes them)
    std::vector<char> skip(nn, 0);
    for (int k = 0; k < flagged[0]; ++k)
        for (int c = 0; c < 32; ++c) {
            const long long col = static_cast<long long>(flagged[1 + k]) * 32 + c - head;
            if (col >= 0 && col < nn) skip[col] = 1;
        }
    // ... nor columns of tiles the serial run (tile 0 at column 0) flagged
    for (int k = 0; k < nf; ++k)
        for (int c = 0; c < 32; ++c) {
            const long long col = static_cast<long long>(fl[k]) * 32 + c - n0;
            if (col >= 0 && col < nn) skip[col] = 1;
        }
    int bad = 0, compared = 0;
    for (int pl = 0; pl < nplanes; ++pl)
        for (long long c = 0; c < nn; ++c) {
            if (skip[c]) continue;
            const float g = out[static_cast<size_t>(pl) * nn + c], e = ref[static_cast<size_t>(pl) * N + n0 + c];
            ++compared;
            if (memcmp(&g, &e, 4) != 0 && !(g != g && e != e)) {
                if (bad < 6) printf("R=%d nw=%d plane %d col %lld: %.9g vs %.9g\n", R, nw, pl, n0 + c, g, e);
                ++bad;
            }
        }
    if (n0 == 0) {                               // same tiling as the serial run: same flagged tiles
        std::vector<int> a(flagged.begin() + 1, flagged.begin() + 1 + flagged[0]), b(fl.begin(), fl.begin() + nf);
        std::sort(a.begin(), a.end());
        std::sort(b.begin(), b.end());
        if (a != b) { printf("flagged tiles differ (%d vs %d)\n", flagged[0], nf); ++bad; }
    }
    if (compared < nplanes * 32) { printf("too few columns compared\n"); ++bad; }
    return bad;
}

int main(int argc, char** argv) {
    const bool quick = argc > 1;
    int bad = 0;
    bad += one(1000, 16, 2, 6, 140, false);    // <16,64>: the reference's call shape, 5 tiles on 2 CTAs
    bad += one(200, 16, 3, 0, 150, false);     // <16,16>
    if (!quick) {
        bad += one(1000, 32, 2, 0, 150, true); // <32,32> with the mode plane
        bad += one(500, 16, 2, 34, 100, true); // <16,32>
        bad += one(100, 32, 2, 0, 150, false); // <32,4>
    }
    if (bad) { printf("%d mismatches\n", bad); return 1; }
    printf("clean\n");
    return 0;
}
