// k2_grid_csort.cu -- K2, counting-sort variant: realization-axis reduction
// with exact order statistics for 32 < R <= 1024.
//
// Same contract as k2_grid_sort.cu (GridFiles.stats cell loop,
// GSIMCLI/tools/grid.py:662-701): per grid node, R realization values ->
// moments, any number of linearly interpolated quantiles, and the mode.
//
// A comparison sort of a 1024-value column costs ~110 instructions per value
// (k2_grid_sort.cu: 11 % of the HBM roofline).  This kernel sorts by
// DISTRIBUTION instead.  One persistent CTA per SM (16 / 24 / 32 warps for
// R <= 1024 / 512 / 256):
//   HBM stack [R][pitch] f32 --TMA 32-node boxes, 128B swizzle--> one tile per
//   CTA; its 32 columns are handed to the warps by a ticket counter; a warp
//   copies its column to registers (lane l holds rows 32 j + l) and the warp
//   that takes the last column re-arms the mbarrier and issues the TMA of the
//   CTA's next tile.  Per column, in registers + a per-warp scratch region:
//   pass 0  nodata/NaN -> +INF, shifted moments (packed f32x2), min / max
//   pass 1  bin = linear map of [min, max] onto B = 32 (IPT+1) bins; values
//           equal to min / max go to per-lane atom bins (DSS clamps, so the
//           extremes are often runs of hundreds of equal values); one
//           shared-memory ATOMS.ADD per value returns its arrival index
//   scan    per-bin exclusive prefix (lane l owns bins l (IPT+1)...: stride
//           IPT+1 keeps every LDS/STS bank-conflict free)
//   pass 2  value -> slot prefix[bin] + arrival (the slot array aliases the
//           histogram: prefixes are read into registers first)
//   big     bins of more than HALO+1 values: skipped when all equal (heavy
//           ties), rank-sorted by the warp otherwise
//   windows lane l re-reads slots [l*IPT - HALO, (l+1)*IPT + HALO) into
//           registers and runs max-bin-size odd-even transposition phases:
//           bins are value-ordered, so exchanges never cross a bin boundary
//           and every bin <= HALO+1 values that touches the lane's own slots
//           is sorted completely inside the window
//   a column whose bin map collapsed (an outlier: one mixed bin of > 128
//   values) takes a generic shared-memory bitonic sort instead -- always exact.
//   finish  quantile lerp in float64 (NumPy _lerp), mode from the equal-
//           neighbour bitmap, moments in float32/64; lane = output plane.
// HBM traffic: the stack is read exactly once (4 B per node-realization).
#include "k2_csort_kernel.cuh"

namespace {

using namespace csort_k;

template <int IPT, bool MODE, int HALO>
int cs_launch_one(gsb_stack* s, const K2Args& args, int grid, cudaStream_t st) {
    constexpr int CS_WARPS = CsWarps<IPT>::v;
    constexpr int NPAD = IPT * 32;
    const size_t smem = static_cast<size_t>(NPAD) * 128
        + static_cast<size_t>(CS_WARPS) * K2_CS_REGION_WORDS(IPT, HALO) * 4 + 16
        + GSB_MAX_PLANES * 12 + CS_WARPS * 100 * 4;
    auto kern = k2_csort_kernel<IPT, MODE, HALO>;
    GSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
    kern<<<grid, CS_WARPS * 32, smem, st>>>(s->tmap_k2, args);
    GSB_CUDA(cudaGetLastError());
    return GSB_OK;
}

template <bool MODE, int HALO>
int cs_launch_ipt(gsb_stack* s, const K2Args& args, int ipt, int grid, cudaStream_t st) {
    switch (ipt) {
    case 2: return cs_launch_one<2, MODE, HALO>(s, args, grid, st);
    case 4: return cs_launch_one<4, MODE, HALO>(s, args, grid, st);
    case 8: return cs_launch_one<8, MODE, HALO>(s, args, grid, st);
    case 16: return cs_launch_one<16, MODE, HALO>(s, args, grid, st);
    case 32: return cs_launch_one<32, MODE, HALO>(s, args, grid, st);
    }
    gsb_set_error("k2 csort: unsupported items-per-lane %d", ipt);
    return GSB_EINVAL;
}

}  // namespace

extern "C" int gsb_debug_cs_counters(unsigned long long* out, int reset) {
    GSB_CUDA(cudaMemcpyFromSymbol(out, gsb_cs_counters, sizeof(unsigned long long) * 8));
    if (reset) {
        unsigned long long z[8] = {0};
        GSB_CUDA(cudaMemcpyToSymbol(gsb_cs_counters, z, sizeof(z)));
    }
    return GSB_OK;
}

int gsb_launch_k2_csort(gsb_stack* s, const K2Args& args, int ipt, int grid, bool mode,
                        int halo, cudaStream_t st) {
    (void)halo;      // one window halo (10) is built: 8 and 12 measured slower
    return mode ? cs_launch_ipt<true, 10>(s, args, ipt, grid, st)
                : cs_launch_ipt<false, 10>(s, args, ipt, grid, st);
}
