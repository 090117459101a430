// k2_rsel_emul.cpp -- TEST INFRASTRUCTURE: runs the per-thread phases of the K2
// register-tile kernel (gsimcli_b200/csrc/k2_rsel_phases.cuh, the very code the
// CUDA kernel inlines) on the host, thread by thread, with the kernel's barriers
// as loop boundaries.  It lets the CPU test suite check the kernel's algorithm
// (bin map, scatter, rank look-up, mode, moments) against the oracle where no
// GPU exists.  Never linked into libgsimcli_b200.so.
//   g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC -o tests/emul/libk2rsel_emul.so tests/emul/k2_rsel_emul.cpp
#include <stdlib.h>
#include <vector>
#include "../../gsimcli_b200/csrc/k2_rsel_phases.cuh"

using namespace rsel;

template <int NW, int IPT>
static int run(const float* data, long long pitch, long long N, int R, float nodata, int nplanes,
               const unsigned char* kind, const double* q, float* out, long long out_pitch,
               int* flagged) {
    constexpr int NT = NW * 32;
    constexpr int S = NW / 8;
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
    int sort_words = R * COLS > 6 * NW * COLS ? R * COLS : 6 * NW * COLS;
    std::vector<float> sortbuf(sort_words);
    std::vector<uint32_t> cells(NCELLS * COLS), colinfo(CI_WORDS * COLS);
    std::vector<double> part2(6 * S * COLS);
    std::vector<float> res(GSB_MAX_PLANES * COLS), pivc(NW * COLS);
    std::vector<unsigned long long> modekey(COLS);
    Smem s;
    s.sortbuf = sortbuf.data();
    s.part = reinterpret_cast<uint32_t*>(sortbuf.data());
    s.chunk = reinterpret_cast<uint32_t*>(sortbuf.data());
    s.cells = cells.data();
    s.part2 = part2.data();
    s.res = res.data();
    s.pivc = pivc.data();
    s.colinfo = colinfo.data();
    s.modekey = modekey.data();
    std::vector<Thr<IPT>> th(NT);
    const long long ntiles = (N + COLS - 1) / COLS;
    int nflag = 0;
    for (long long tile = 0; tile < ntiles; ++tile) {
        for (int tid = 0; tid < NT; ++tid) {
            const int w = tid >> 5, c = tid & 31;
            const long long col = tile * COLS + c;
            for (int i = 0; i < IPT; ++i) {
                const int r = i * NW + w;
                th[tid].xs[i] = (col < N && r < R) ? data[(long long)r * pitch + col] : qnan();
            }
        }
        for (auto& v : cells) v = 0;
#define ALL(stmt) for (int tid = 0; tid < NT; ++tid) { const int w = tid >> 5, c = tid & 31; (void)w; (void)c; stmt; }
        ALL((phase_pivot_write<NW, IPT>(th[tid], s, a, w, c)))
        ALL((phase_pivot_read<NW, IPT>(th[tid], s, c)); (phase_a<NW, IPT>(th[tid], s, a, w, c)))
        ALL((phase_combine1<NW>(s, w, c)))
        ALL((phase_combine2<NW, IPT>(th[tid], s, w, c)); (phase_b<NW, IPT>(th[tid], s, c)))
        ALL((phase_scan1<NW>(s, w, c)))
        ALL((phase_scan2<NW>(s, w, c)))
        ALL((phase_c<NW, IPT>(th[tid], s, c)))
        ALL((phase_fin_quant<NW, IPT>(th[tid], s, a, w, c)); (phase_fin_mode<NW, IPT>(th[tid], s, w, c)))
        bool fl = false;
        for (int c = 0; c < COLS; ++c) fl = fl || colinfo[CI_FLAG * COLS + c] != 0u;
        if (fl) { if (flagged) flagged[nflag] = (int)tile; ++nflag; }
        for (int tid = 0; tid < NT; ++tid) {
            const int w = tid >> 5, c = tid & 31;
            const long long col = tile * COLS + c;
            for (int pl = w; pl < nplanes; pl += NW) {
                const float v = plane_value<NW, IPT>(th[tid], s, a, pl, c);
                if (col < N) out[(long long)pl * out_pitch + col] = v;
            }
        }
    }
    return nflag;
}

// returns the number of flagged tiles (their planes are NOT valid: the CUDA path
// hands them to the counting-sort kernel), or -1 for an unsupported R
extern "C" int k2_rsel_emulate(const float* data, long long pitch, long long N, int R, float nodata,
                               int nplanes, const unsigned char* kind, const double* q, float* out,
                               long long out_pitch, int nw, int* flagged) {
    if (R < 1 || R > 1024) return -1;
    if (nodata != nodata) nodata = INFINITY;
    if (nw == 16) {
        const int ipt = (R + 15) / 16;
        if (ipt <= 4) return run<16, 4>(data, pitch, N, R, nodata, nplanes, kind, q, out, out_pitch, flagged);
        if (ipt <= 8) return run<16, 8>(data, pitch, N, R, nodata, nplanes, kind, q, out, out_pitch, flagged);
        if (ipt <= 16) return run<16, 16>(data, pitch, N, R, nodata, nplanes, kind, q, out, out_pitch, flagged);
        if (ipt <= 32) return run<16, 32>(data, pitch, N, R, nodata, nplanes, kind, q, out, out_pitch, flagged);
        return run<16, 64>(data, pitch, N, R, nodata, nplanes, kind, q, out, out_pitch, flagged);
    }
    const int ipt = (R + 31) / 32;
    if (ipt <= 2) return run<32, 2>(data, pitch, N, R, nodata, nplanes, kind, q, out, out_pitch, flagged);
    if (ipt <= 4) return run<32, 4>(data, pitch, N, R, nodata, nplanes, kind, q, out, out_pitch, flagged);
    if (ipt <= 8) return run<32, 8>(data, pitch, N, R, nodata, nplanes, kind, q, out, out_pitch, flagged);
    if (ipt <= 16) return run<32, 16>(data, pitch, N, R, nodata, nplanes, kind, q, out, out_pitch, flagged);
    return run<32, 32>(data, pitch, N, R, nodata, nplanes, kind, q, out, out_pitch, flagged);
}
