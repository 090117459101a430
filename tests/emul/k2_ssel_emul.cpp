// k2_ssel_emul.cpp -- TEST INFRASTRUCTURE: runs the per-thread phases of the K2
// staged-select kernel (gsimcli_b200/csrc/k2_ssel_phases.cuh, the very code the
// CUDA kernel inlines) on the host, thread by thread, with the kernel's barriers
// as loop boundaries and the cp.async tile copy as a plain copy into the stage.
// Never linked into libgsimcli_b200.so.
//   g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC -o tests/emul/libk2ssel_emul.so tests/emul/k2_ssel_emul.cpp
#include <stdlib.h>
#include <vector>
#include "../../gsimcli_b200/csrc/k2_ssel_phases.cuh"

using namespace ssel;

template <int NW, int IPT>
static int run(const float* data, long long pitch, long long N, int R, float nodata, int nplanes,
               const unsigned char* kind, const double* q, float* out, long long out_pitch,
               int* flagged) {
    constexpr int NT = NW * 32;
    constexpr int ROWS = NW * IPT;
    Args a;
    memset(&a, 0, sizeof(a));
    a.R = R;
    a.nplanes = nplanes;
    a.nodata = nodata;
    int nq = 0;
    for (int i = 0; i < nplanes; ++i) {
        a.kind[i] = kind[i];
        a.q[i] = q[i];
        if (kind[i] == PK_QUANT) a.qplane[nq++] = (unsigned char)i;
    }
    a.nq = nq;
    if (nq < 1 || 2 * nq > NLMAX) return -1;
    // exact sizes: an out-of-bounds index of a phase is an ASan / valgrind error
    std::vector<float> stage(ROWS * COLS, qnan());
    std::vector<uint32_t> cells(NCELLS * COLS, 0u);
    const int list_words = NLMAX * LSTRIDE, part_words = 3 * NW * COLS;
    std::vector<uint32_t> lists(list_words > part_words ? list_words : part_words, 0u);
    std::vector<uint32_t> rinfo(NLMAX * COLS), chunk(NW * COLS), mnmx(2 * COLS), ncol(COLS);
    std::vector<double> msum(3 * COLS);
    Smem s;
    s.stage = stage.data();
    s.cells = cells.data();
    s.lists = lists.data();
    s.part = lists.data();           // same alias as the kernel
    s.rinfo = rinfo.data();
    s.chunk = chunk.data();
    s.mnmx = mnmx.data();
    s.msum = msum.data();
    s.ncol = ncol.data();
    std::vector<Thr<IPT>> th(NT);
    const long long ntiles = (N + COLS - 1) / COLS;
    int nflag = 0;
    for (long long tile = 0; tile < ntiles; ++tile) {
        // the tile copy: rows < R, 16-byte chunks whose start is clamped into the row
        for (int r = 0; r < R; ++r)
            for (int ch = 0; ch < 8; ++ch) {
                long long col4 = tile * COLS + ch * 4;
                if (col4 > pitch - 4) col4 = pitch - 4;
                for (int k = 0; k < 4; ++k) stage[r * COLS + ch * 4 + k] = data[(long long)r * pitch + col4 + k];
            }
        for (int c = 0; c < COLS; ++c) { mnmx[c] = 0xffffffffu; mnmx[COLS + c] = 0u; }
#define ALL(stmt) for (int tid = 0; tid < NT; ++tid) { const int w = tid >> 5, c = tid & 31; (void)w; (void)c; stmt; }
        ALL((phase_a<NW, IPT>(th[tid], s, a, w, c)))
        // (the kernel issues the copy of the next tile here)
        for (auto& v : stage) v = u2f(0xdeadbeefu);      // the stage is dead from here on
        ALL((phase_c1<NW, IPT>(th[tid], s, w, c)); (phase_b<NW, IPT>(th[tid], s, c)))
        ALL((phase_s1<NW>(s, w, c)))
        ALL((phase_locate<NW, IPT>(th[tid], s, a, w, c)))
        ALL((phase_x<NW, IPT>(th[tid], s, c)))
        bool fl = false;
        for (int tid = 0; tid < NT; ++tid) {
            const int w = tid >> 5, c = tid & 31;
            const long long col = tile * COLS + c;
            bool f = th[tid].dead != 0;
            for (int pl = w; pl < nplanes; pl += NW) {
                const float v = plane_value<NW, IPT>(th[tid], s, a, pl, c, &f);
                if (col < N) out[(long long)pl * out_pitch + col] = v;
            }
            if (col < N && f) fl = true;
        }
        for (auto& v : cells) v = 0u;
        for (int r = R; r < ROWS; ++r)
            for (int c = 0; c < COLS; ++c) stage[r * COLS + c] = qnan();
        if (fl) { if (flagged) flagged[nflag] = (int)tile; ++nflag; }
    }
    return nflag;
}

// returns the number of flagged tiles (their planes are NOT valid: the CUDA path
// hands them to the counting-sort kernel), or -1 for an unsupported shape
extern "C" int k2_ssel_emulate(const float* data, long long pitch, long long N, int R, float nodata,
                               int nplanes, const unsigned char* kind, const double* q, float* out,
                               long long out_pitch, int nw, int* flagged) {
    if (R < 1 || R > 1024 || pitch < 4 || (pitch & 3)) return -1;
    if (nodata != nodata) nodata = INFINITY;
    if (nw == 16) {
        const int ipt = (R + 15) / 16;
        if (ipt <= 16) return run<16, 16>(data, pitch, N, R, nodata, nplanes, kind, q, out, out_pitch, flagged);
        if (ipt <= 32) return run<16, 32>(data, pitch, N, R, nodata, nplanes, kind, q, out, out_pitch, flagged);
        return run<16, 64>(data, pitch, N, R, nodata, nplanes, kind, q, out, out_pitch, flagged);
    }
    const int ipt = (R + 31) / 32;
    if (ipt <= 8) return run<32, 8>(data, pitch, N, R, nodata, nplanes, kind, q, out, out_pitch, flagged);
    if (ipt <= 16) return run<32, 16>(data, pitch, N, R, nodata, nplanes, kind, q, out, out_pitch, flagged);
    return run<32, 32>(data, pitch, N, R, nodata, nplanes, kind, q, out, out_pitch, flagged);
}
