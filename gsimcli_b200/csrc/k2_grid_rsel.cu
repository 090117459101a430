// k2_grid_rsel.cu -- K2, "register tile" kernel: realization-axis reduction with
// exact order statistics for 32 < R <= 1024 (GridFiles.stats cell loop,
// GSIMCLI/tools/grid.py:662-701).
//
// Why this shape (profiles/ubench/r02_ubench2_b200.log):
//   * a [R x 32 nodes] tile is the narrowest that streams at full rate (4.6 TB/s
//     through one shared-memory buffer per SM; 16-node = 64-byte rows stop at
//     2.3 TB/s whatever the occupancy), and 125 KB of it fit only once per SM:
//     a shared-memory tile cannot be double buffered, and refilling a single
//     buffer leaves 4 us per tile exposed;
//   * so the tile lives in REGISTERS: thread (warp w, lane c) keeps rows
//     w, w+NW, ... of column c.  The next tile is fetched into those registers
//     by coalesced 128-byte row loads as soon as the scatter of the current one
//     is done, and lands while the finish phase runs (125 KB in flight per SM);
//   * lane = column makes every shared-memory access of the histogram, the
//     scatter and the finish phase bank-conflict free (1.0 cycle per warp
//     instruction instead of 3.4 for a warp scattering one column).
// The per-thread phases are in k2_rsel_phases.cuh (shared with the host
// emulation in tests/emul); this file is the tile loop, the barriers and the
// launcher.  Tiles the kernel cannot finish exactly (flag) are listed in global
// memory and redone by the counting-sort kernel on the same stream.
// Production use (k2_grid_sort.cu): the reference's call shape (<= 3 quantile
// planes, no mode) at R >= 200, where it is 10-20 % faster than the counting
// sort; with many quantile planes it ties and with the mode plane it is 2.3x
// slower (lane = column turns per-column decisions into warp divergence), so
// those stay with the counting sort.  DESIGN.md sections 4 and 8.
#include "gsb_common.cuh"
#include "k2_args.cuh"
#include "k2_rsel_kernel.cuh"

namespace {

using namespace rsel;

template <int NW, int IPT>
size_t rsel_smem(int R, int* sort_words) {
    const int need_rows = R * COLS;
    const int need_alias = 6 * NW * COLS;       // `part` is the largest alias
    const int words = ((need_rows > need_alias ? need_rows : need_alias) + 3) & ~3;
    *sort_words = words;
    const size_t fixed = (static_cast<size_t>(NCELLS) * COLS * 4 + 6 * (NW / 8) * COLS * 8 + COLS * 8 +
                          GSB_MAX_PLANES * COLS * 4 + NW * COLS * 4 + CI_WORDS * COLS * 4 + 15) & ~size_t(15);
    return fixed + static_cast<size_t>(words) * 4;
}

template <int NW, int IPT, bool MODE>
int rsel_launch_one(RselLaunch& L, int grid, cudaStream_t st) {
    const size_t smem = rsel_smem<NW, IPT>(L.a.R, &L.sort_words);
    auto kern = k2_rsel_kernel<NW, IPT, MODE>;
    GSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, NW * 32, smem, st>>>(L);
    GSB_CUDA(cudaGetLastError());
    return GSB_OK;
}

template <int NW, bool MODE>
int rsel_launch_nw(RselLaunch& L, int grid, cudaStream_t st) {
    const int ipt = (L.a.R + NW - 1) / NW;
    // IPT = smallest power of two >= ipt, so that rows below IPT/2 * NW < R always exist
    if constexpr (NW == 32) {
        if (ipt <= 2) return rsel_launch_one<NW, 2, MODE>(L, grid, st);
    }
    if (ipt <= 4) return rsel_launch_one<NW, 4, MODE>(L, grid, st);
    if (ipt <= 8) return rsel_launch_one<NW, 8, MODE>(L, grid, st);
    if (ipt <= 16) return rsel_launch_one<NW, 16, MODE>(L, grid, st);
    if (ipt <= 32) return rsel_launch_one<NW, 32, MODE>(L, grid, st);
    if constexpr (NW <= 16) {
        if (ipt <= 64) return rsel_launch_one<NW, 64, MODE>(L, grid, st);
    }
    gsb_set_error("k2 rsel: R = %d does not fit %d warps", L.a.R, NW);
    return GSB_EINVAL;
}

}  // namespace

// K2Args carries the plane list and the tile geometry (k2_grid_sort.cu fills
// it).  `nw` = warps per CTA: 32 (32 values per thread at R = 1024) or 16.
int gsb_launch_k2_rsel(gsb_stack* s, const K2Args& args, int grid, bool mode, int nw, int* d_flagged,
                       cudaStream_t st) {
    RselLaunch L;
    memset(&L, 0, sizeof(L));
    L.a.R = args.R;
    L.a.nplanes = args.nplanes;
    // an ordered compare never matches NaN: +INF stands in (the tile kernels
    // treat +INF as missing anyway)
    L.a.nodata = (args.nodata == args.nodata) ? args.nodata : INFINITY;
    int nq = 0;
    for (int i = 0; i < args.nplanes; ++i) {
        L.a.kind[i] = args.kind[i];
        L.a.q[i] = args.q[i];
        if (args.kind[i] == GSB_PK_QUANT) L.a.qplane[nq++] = static_cast<unsigned char>(i);
    }
    L.a.nq = nq;
    L.data = s->data;
    L.pitch = s->pitch;
    L.N = s->N;
    L.ntiles = args.ntiles;
    L.n0 = args.n0;
    L.nn = args.nn;
    L.skip = args.skip;
    L.out = args.out;
    L.out_pitch = args.out_pitch;
    L.flagged = d_flagged;
    if (nw == 16)
        return mode ? rsel_launch_nw<16, true>(L, grid, st) : rsel_launch_nw<16, false>(L, grid, st);
    return mode ? rsel_launch_nw<32, true>(L, grid, st) : rsel_launch_nw<32, false>(L, grid, st);
}
