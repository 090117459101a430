// k2_grid_ssel.cu -- K2, "staged select" kernel: realization-axis reduction for
// the reference's own call shape (GridFiles.stats(lmean, lmed, lskew, lperc),
// GSIMCLI/tools/grid.py:601-602, cell loop :662-701): moment planes + at most 3
// quantile planes, no mode, 128 < R <= 1024.
//
// EXPERIMENTAL, opt-in (gsb_stack_set_option "k2_algo" = 4): it was written after
// the round's GPU budget was spent and has never run on a device at commit time,
// so the default dispatch does not use it.  bench.py times it in a subprocess and
// checks it bit for bit against the default kernel; tests/emul runs its phases on
// the CPU.  DESIGN.md section 8 has the reasoning.
//
// One persistent CTA per SM, tile = 32 adjacent nodes x R realizations, LANE =
// COLUMN (k2_ssel_phases.cuh).  The tile being reduced lives in registers; the
// NEXT tile is copied global -> shared memory by cp.async (16-byte chunks, 128-
// byte row segments, L1 bypassed) as soon as the stage has been read, so the copy
// engine, not the warps, waits for HBM -- the register-tile kernel (k2_grid_rsel.cu)
// can only prefetch during its finish phase, which the reference's call shape
// makes too short (24 % of its stall samples are that wait).  The 125 KB sort
// buffer of that kernel is what the stage replaces: nothing is scattered, the
// wanted ranks are selected from the histogram (phases L, X, O).
// Every loop is bounded and every barrier is reached by all threads of the CTA
// (no spin waits: cp.async.wait_group + __syncthreads only).
#include "gsb_common.cuh"
#include "k2_args.cuh"
#include "k2_ssel_kernel.cuh"

namespace {

using namespace ssel;

template <int NW, int IPT>
int ssel_launch_one(const SselLaunch& L, int grid, cudaStream_t st) {
    constexpr int smem = Layout<NW, IPT>::BYTES;
    auto kern = k2_ssel_kernel<NW, IPT>;
    GSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, NW * 32, smem, st>>>(L);
    GSB_CUDA(cudaGetLastError());
    return GSB_OK;
}

}  // namespace

// true when the staged-select kernel can serve the call: at most NLMAX / 2
// quantile planes and at least one, no mode plane, a deep column
bool gsb_k2_ssel_supports(const K2Args& args, bool mode) {
    int nq = 0;
    for (int i = 0; i < args.nplanes; ++i) nq += args.kind[i] == GSB_PK_QUANT ? 1 : 0;
    return !mode && nq >= 1 && 2 * nq <= NLMAX && args.R > 128 && args.R <= 1024;
}

// K2Args as for the register-tile kernel (k2_grid_sort.cu fills it); `nw` = warps
// per CTA: 16 (up to 64 values per thread) or 32 (up to 32).  Tiles that cannot
// be finished exactly are appended to d_flagged for the counting-sort kernel.
int gsb_launch_k2_ssel(gsb_stack* s, const K2Args& args, int grid, bool mode, int nw, int* d_flagged,
                       cudaStream_t st) {
    GSB_REQUIRE(gsb_k2_ssel_supports(args, mode), "k2 ssel: call shape not supported");
    GSB_REQUIRE(s->pitch >= 32 && (s->pitch & 31) == 0 && (args.n0 & 3) == 0, "k2 ssel: unaligned stack");
    SselLaunch L;
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
    L.ntiles = args.ntiles;
    L.n0 = args.n0;
    L.nn = args.nn;
    L.skip = args.skip;
    L.out = args.out;
    L.out_pitch = args.out_pitch;
    L.flagged = d_flagged;
    const int R = args.R;
    if (nw == 32) {
        if (R <= 256) return ssel_launch_one<32, 8>(L, grid, st);
        if (R <= 512) return ssel_launch_one<32, 16>(L, grid, st);
        return ssel_launch_one<32, 32>(L, grid, st);
    }
    if (R <= 256) return ssel_launch_one<16, 16>(L, grid, st);
    if (R <= 512) return ssel_launch_one<16, 32>(L, grid, st);
    return ssel_launch_one<16, 64>(L, grid, st);
}
