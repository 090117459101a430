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
#include "k2_ssel_phases.cuh"

namespace {

using namespace ssel;

struct SselLaunch {
    Args a;
    const float* data;      // row 0, node 0 of the stack
    long long pitch;
    long long ntiles;
    long long n0;           // first column of tile 0 (multiple of 4)
    long long nn;           // columns from n0 to the end of the requested range
    long long skip;         // leading columns of tile 0 before the requested range
    float* out;             // plane 0, column n0
    long long out_pitch;
    int* flagged;           // [0] count, [1 + k] tile index
};

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// rows 0..R-1 of the tile -> stage[row][32]: thread `tid` copies chunk tid & 7
// (4 nodes) of rows tid >> 3, + NT / 8, ...  A chunk that would start past the
// end of the row (last tile, or a sub-range that is not 32-aligned) re-reads the
// row's last chunk: always inside the allocation, never stored.
template <int NT>
__device__ __forceinline__ void issue_tile(const SselLaunch& L, float* stage, long long tile, int tid) {
    const int ch = tid & 7;
    long long col4 = L.n0 + tile * COLS + ch * 4;
    col4 = col4 > L.pitch - 4 ? L.pitch - 4 : col4;
    const int R = L.a.R;
    int row = tid >> 3;
    const float* src = L.data + static_cast<long long>(row) * L.pitch + col4;
    uint32_t dst = gsb_smem_u32(stage + row * COLS + ch * 4);
    const long long sstep = static_cast<long long>(NT / 8) * L.pitch;
    for (; row < R; row += NT / 8) {
        cp_async16(dst, src);
        src += sstep;
        dst += (NT / 8) * COLS * 4;
    }
    cp_async_commit();
}

template <int NW, int IPT>
struct Layout {
    static constexpr int ROWS = NW * IPT;
    static constexpr int LIST_WORDS = NLMAX * LSTRIDE;
    static constexpr int PART_WORDS = 3 * NW * COLS;
    static constexpr int OFF_CELLS = 0;
    static constexpr int OFF_LISTS = OFF_CELLS + NCELLS * COLS * 4;
    static constexpr int OFF_RINFO = OFF_LISTS + (LIST_WORDS > PART_WORDS ? LIST_WORDS : PART_WORDS) * 4;
    static constexpr int OFF_CHUNK = OFF_RINFO + NLMAX * COLS * 4;
    static constexpr int OFF_MNMX = OFF_CHUNK + NW * COLS * 4;
    static constexpr int OFF_NCOL = OFF_MNMX + 2 * COLS * 4;
    static constexpr int OFF_MSUM = (OFF_NCOL + COLS * 4 + 7) & ~7;
    static constexpr int OFF_STAGE = (OFF_MSUM + 3 * COLS * 8 + 127) & ~127;
    static constexpr int BYTES = OFF_STAGE + ROWS * COLS * 4;
};

template <int NW, int IPT>
__global__ void __launch_bounds__(NW * 32, 1)
k2_ssel_kernel(const __grid_constant__ SselLaunch L) {
    constexpr int NT = NW * 32;
    using LO = Layout<NW, IPT>;
    static_assert(LO::BYTES <= 227 * 1024, "shared memory layout too large");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem s;
    s.cells = reinterpret_cast<uint32_t*>(smem_raw + LO::OFF_CELLS);
    s.lists = reinterpret_cast<uint32_t*>(smem_raw + LO::OFF_LISTS);
    s.part = s.lists;
    s.rinfo = reinterpret_cast<uint32_t*>(smem_raw + LO::OFF_RINFO);
    s.chunk = reinterpret_cast<uint32_t*>(smem_raw + LO::OFF_CHUNK);
    s.mnmx = reinterpret_cast<uint32_t*>(smem_raw + LO::OFF_MNMX);
    s.ncol = reinterpret_cast<uint32_t*>(smem_raw + LO::OFF_NCOL);
    s.msum = reinterpret_cast<double*>(smem_raw + LO::OFF_MSUM);
    s.stage = reinterpret_cast<float*>(smem_raw + LO::OFF_STAGE);
    const int tid = threadIdx.x;
    const int w = tid >> 5, c = tid & 31;
    const Args& a = L.a;

    long long tile = blockIdx.x;
    if (tile >= L.ntiles) return;                    // whole CTA: before any barrier
    issue_tile<NT>(L, s.stage, tile, tid);
    // rows R .. ROWS-1 of the stage are never copied: NaN once (missing values)
    for (int i = a.R * COLS + tid; i < LO::ROWS * COLS; i += NT) s.stage[i] = qnan();
    {
        uint4* z = reinterpret_cast<uint4*>(s.cells);
        for (int i = tid; i < NCELLS * COLS / 4; i += NT) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (w == 0) {
        s.mnmx[c] = 0xffffffffu;
        s.mnmx[COLS + c] = 0u;
    }
    cp_async_wait_all();
    __syncthreads();

    Thr<IPT> t;
    for (;;) {
        phase_a<NW, IPT>(t, s, a, w, c);
        __syncthreads();
        // every thread has read its rows: the stage takes the next tile, which
        // lands while this one is reduced
        const long long next = tile + gridDim.x;
        const bool has_next = next < L.ntiles;       // CTA-uniform
        if (has_next) issue_tile<NT>(L, s.stage, next, tid);
        phase_c1<NW, IPT>(t, s, w, c);
        phase_b<NW, IPT>(t, s, c);
        __syncthreads();
        phase_s1<NW>(s, w, c);
        __syncthreads();
        phase_locate<NW, IPT>(t, s, a, w, c);
        __syncthreads();
        phase_x<NW, IPT>(t, s, c);
        __syncthreads();
        // planes of the finished tile; lane = column -> 128-byte row stores
        bool flag = t.dead != 0;
        {
            const long long col = tile * COLS + c;
            const bool colok = col >= L.skip && col < L.nn;
            for (int pl = w; pl < a.nplanes; pl += NW) {          // warp-uniform
                const float v = plane_value<NW, IPT>(t, s, a, pl, c, &flag);
                if (colok) L.out[static_cast<long long>(pl) * L.out_pitch + col] = v;
            }
            flag = flag && colok;
            uint4* z = reinterpret_cast<uint4*>(s.cells);
            for (int i = tid; i < NCELLS * COLS / 4; i += NT) z[i] = make_uint4(0u, 0u, 0u, 0u);
            if (w == NW - 1) {
                s.mnmx[c] = 0xffffffffu;
                s.mnmx[COLS + c] = 0u;
            }
        }
        cp_async_wait_all();
        const int any = __syncthreads_or(flag ? 1 : 0);
        if (any && tid == 0) {
            const int k = atomicAdd(L.flagged, 1);
            L.flagged[1 + k] = static_cast<int>(tile);
        }
        if (!has_next) break;
        tile = next;
    }
}

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
