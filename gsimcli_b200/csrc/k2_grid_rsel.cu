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
#include "k2_rsel_phases.cuh"

namespace {

using namespace rsel;

struct RselLaunch {
    Args a;
    const float* data;      // row 0, node 0 of the stack
    long long pitch;
    long long N;            // nodes held by the stack
    long long ntiles;
    long long n0;           // first column of tile 0 (multiple of 4)
    long long nn;           // columns from n0 to the end of the requested range
    long long skip;         // leading columns of tile 0 before the requested range
    float* out;             // plane 0, column n0
    long long out_pitch;
    int* flagged;           // [0] count, [1 + k] tile index
    int sort_words;         // words of the sortbuf region (>= R * 32 and the aliases)
};

__device__ __forceinline__ float ld_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// rows w, w + NW, ... of column (tile, c): one coalesced 128-byte row segment per
// warp instruction, no branches, nothing that waits for the data (the loads of
// the next tile stay in flight while the current one is finished).  A column
// past the end of the stack re-reads the last column (its planes are never
// stored); a row past R keeps the NaN its register was preset with (only the
// upper half of the row groups can be past R: IPT/2 * NW < R by construction).
__device__ __forceinline__ float ld_stream_if(const float* p, bool ok) {
    float v = __int_as_float(0x7fc00000);
    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p ld.global.nc.L1::no_allocate.f32 %0, [%1];\n}"
                 : "+f"(v) : "l"(p), "r"(static_cast<unsigned>(ok)));
    return v;
}

template <int NW, int IPT>
__device__ __forceinline__ void load_tile(Thr<IPT>& t, const RselLaunch& L, long long tile, int w, int c) {
    const long long col = L.n0 + tile * COLS + c;
    const int R = L.a.R;
    const float* p = L.data + (col < L.N ? col : L.N - 1) + static_cast<long long>(w) * L.pitch;
    const long long step = static_cast<long long>(NW) * L.pitch;      // one row group further
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        if (i < IPT / 2) t.xs[i] = ld_stream(p);
        else t.xs[i] = ld_stream_if(p, i * NW + w < R);
        p += step;
    }
}

template <int NW, int IPT, bool MODE>
__global__ void __launch_bounds__(NW * 32, 1)
k2_rsel_kernel(const __grid_constant__ RselLaunch L) {
    constexpr int NT = NW * 32;
    constexpr int S = NW / 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // fixed-size regions first, at compile-time offsets (their addresses fold
    // into the instructions); the sort buffer, whose size follows R, is last
    constexpr int OFF_PART2 = NCELLS * COLS * 4;
    constexpr int OFF_MODE = OFF_PART2 + 6 * S * COLS * 8;
    constexpr int OFF_RES = OFF_MODE + COLS * 8;
    constexpr int OFF_PIVC = OFF_RES + GSB_MAX_PLANES * COLS * 4;
    constexpr int OFF_INFO = OFF_PIVC + NW * COLS * 4;
    constexpr int OFF_SORT = (OFF_INFO + CI_WORDS * COLS * 4 + 15) & ~15;
    Smem s;
    s.cells = reinterpret_cast<uint32_t*>(smem_raw);
    s.part2 = reinterpret_cast<double*>(smem_raw + OFF_PART2);
    s.modekey = reinterpret_cast<unsigned long long*>(smem_raw + OFF_MODE);
    s.res = reinterpret_cast<float*>(smem_raw + OFF_RES);
    s.pivc = reinterpret_cast<float*>(smem_raw + OFF_PIVC);
    s.colinfo = reinterpret_cast<uint32_t*>(smem_raw + OFF_INFO);
    s.sortbuf = reinterpret_cast<float*>(smem_raw + OFF_SORT);
    s.part = reinterpret_cast<uint32_t*>(smem_raw + OFF_SORT);
    s.chunk = reinterpret_cast<uint32_t*>(smem_raw + OFF_SORT);
    const int tid = threadIdx.x;
    const int w = tid >> 5, c = tid & 31;
    const Args& a = L.a;

    long long tile = blockIdx.x;
    if (tile >= L.ntiles) return;
    Thr<IPT> t;
    load_tile<NW, IPT>(t, L, tile, w, c);
    {
        uint4* z = reinterpret_cast<uint4*>(s.cells);
        for (int i = tid; i < NCELLS * COLS / 4; i += NT) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    phase_pivot_write<NW, IPT>(t, s, a, w, c);
    __syncthreads();

    for (;;) {
        phase_pivot_read<NW, IPT>(t, s, c);
        phase_a<NW, IPT>(t, s, a, w, c);
        __syncthreads();
        phase_combine1<NW>(s, w, c);
        __syncthreads();
        phase_combine2<NW, IPT>(t, s, w, c);
        phase_b<NW, IPT>(t, s, c);
        __syncthreads();
        phase_scan1<NW>(s, w, c);
        __syncthreads();
        phase_scan2<NW>(s, w, c);
        __syncthreads();
        phase_c<NW, IPT>(t, s, c);
        __syncthreads();
        // the values are in shared memory now: the registers take the next tile,
        // which lands while this one is finished
        const long long next = tile + gridDim.x;
        const bool has_next = next < L.ntiles;
        if (has_next) load_tile<NW, IPT>(t, L, next, w, c);
        phase_fin_quant<NW, IPT>(t, s, a, w, c);
        if constexpr (MODE) phase_fin_mode<NW, IPT>(t, s, w, c);
        if (has_next) phase_pivot_write<NW, IPT>(t, s, a, w, c);
        __syncthreads();
        // planes of the finished tile; lane = column -> 128-byte row stores
        {
            const long long col = tile * COLS + c;
            const bool colok = col >= L.skip && col < L.nn;
            for (int pl = w; pl < a.nplanes; pl += NW) {
                const float v = plane_value<NW, IPT>(t, s, a, pl, c);
                if (colok) L.out[static_cast<long long>(pl) * L.out_pitch + col] = v;
            }
            if (w == 0) {
                const bool fl = s.colinfo[CI_FLAG * COLS + c] != 0u;
                if (__any_sync(0xffffffffu, fl) && c == 0) {
                    const int k = atomicAdd(L.flagged, 1);
                    L.flagged[1 + k] = static_cast<int>(tile);
                }
            }
            uint4* z = reinterpret_cast<uint4*>(s.cells);
            for (int i = tid; i < NCELLS * COLS / 4; i += NT) z[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        if (!has_next) break;
        tile = next;
    }
}

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
