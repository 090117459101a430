// k2_ssel_tsan_main.cpp -- TEST INFRASTRUCTURE: stand-alone driver of the threaded
// emulation of the staged-select kernel (k2_ssel_threads.cpp), built with
// -fsanitize=thread (or address): the kernel function runs with one OS thread per
// CUDA thread, in both cp.async emulation modes, on a sub-range that is not
// tile-aligned and with more tiles than blocks, and its planes must equal the
// serial emulation's (k2_ssel_emul.cpp) bit for bit.  Exit code 0 and the word
// "clean" on success; ThreadSanitizer reports make the exit code non-zero.
#include <stdio.h>
#include "k2_ssel_emul.cpp"
#include "k2_ssel_threads.cpp"

static unsigned long long rng_state = 0x9E3779B97F4A7C15ull;
static unsigned rnd() {
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return (unsigned)(rng_state >> 32);
}

static int one(int R, int nw, int grid, long long n0, long long nn) {
    const long long N = 200, pitch = 224;
    const float nodata = -999.9f;
    std::vector<float> data((size_t)R * pitch, 1.0e30f);
    for (int r = 0; r < R; ++r)
        for (long long c = 0; c < N; ++c) {
            const unsigned u = rnd();
            float v = 400.0f + (float)(u % 8000) / 16.0f + (float)(c % 7) * 20.0f;
            if (v > 800.0f) v = 800.0f;                       // clamp atom
            if ((u >> 20) % 97 == 0) v = nodata;
            data[(size_t)r * pitch + c] = v;
        }
    for (int r = 0; r < R; ++r) {
        data[(size_t)r * pitch + 11] = 512.25f;                // constant column
        data[(size_t)r * pitch + 12] = nodata;                 // empty column
        if (r < R / 3) data[(size_t)r * pitch + 40] = nodata;  // first rows missing
    }
    const int nplanes = 5;
    const unsigned char kind[5] = {0, 1, 2, 1, 1};
    const double q[5] = {0, 0.5, 0, 0.025, 0.975};
    std::vector<float> ref((size_t)nplanes * N, -1.f);
    std::vector<int> fl(64, 0);
    const int nf = k2_ssel_emulate(data.data(), pitch, N, R, nodata, nplanes, kind, q, ref.data(), N, nw, fl.data());
    if (nf != 0) { printf("serial emulation flagged %d tiles\n", nf); return 1; }
    int bad = 0;
    for (int mode = 0; mode < 2; ++mode) {
        std::vector<float> out((size_t)nplanes * nn, -2.f);
        std::vector<int> flagged(64, 0);
        if (k2_ssel_threads(data.data(), pitch, R, nodata, n0, nn, nplanes, kind, q, out.data(), nn, nw, grid,
                            mode, flagged.data()) != 0) { printf("unsupported\n"); return 1; }
        if (flagged[0] != 0) { printf("threads flagged %d tiles\n", flagged[0]); ++bad; }
        for (int pl = 0; pl < nplanes; ++pl)
            for (long long c = 0; c < nn; ++c) {
                unsigned a, b;
                memcpy(&a, &out[(size_t)pl * nn + c], 4);
                memcpy(&b, &ref[(size_t)pl * N + n0 + c], 4);
                const bool nan_both = out[(size_t)pl * nn + c] != out[(size_t)pl * nn + c] &&
                                      ref[(size_t)pl * N + n0 + c] != ref[(size_t)pl * N + n0 + c];
                if (a != b && !nan_both) {
                    if (bad < 5) printf("R=%d nw=%d mode=%d plane %d col %lld: %g vs %g\n", R, nw, mode, pl, n0 + c,
                                        out[(size_t)pl * nn + c], ref[(size_t)pl * N + n0 + c]);
                    ++bad;
                }
            }
    }
    return bad;
}

int main(int argc, char** argv) {
    const bool quick = argc > 1;          // under ThreadSanitizer: the small shapes only
    int bad = 0;
    bad += one(200, 16, 2, 6, 150);       // <16,16>: 5 tiles on 2 blocks, head = 2, ragged end
    bad += one(200, 16, 3, 0, 200);
    if (!quick) {
        bad += one(1000, 16, 2, 34, 100); // <16,64>
        bad += one(500, 32, 2, 6, 150);   // <32,16>
        bad += one(1000, 32, 3, 0, 200);  // <32,32>
    }
    if (bad) { printf("%d mismatches\n", bad); return 1; }
    printf("clean\n");
    return 0;
}
