// k2_ssel_kernel.cuh -- the tile loop of the K2 staged-select kernel: barriers,
// the cp.async staging of the next tile, the shared-memory layout, the output
// stores.  Kept apart from the launcher (k2_grid_ssel.cu) so that tests/emul can
// compile THIS code for the host and run it with one OS thread per CUDA thread
// and a real barrier (tests/emul/k2_ssel_threads.cpp, also under ThreadSanitizer):
// a missing barrier or a stage that is refilled too early is then a data race on
// the CPU, where the serial emulation of the phases cannot see it.
//
// Host build: SSEL_HOST is defined by the includer, which supplies threadIdx /
// blockIdx / gridDim, __syncthreads, __syncthreads_or, atomicAdd(int*), the
// shared-memory block (ssel_host_smem) and the three copy hooks below.
#pragma once
#include "k2_ssel_phases.cuh"

#ifndef SSEL_HOST
#define SSEL_GLOBAL(NT) __global__ void __launch_bounds__(NT, 1)
#define SSEL_PARAM const __grid_constant__
#define SSEL_DEV __device__ __forceinline__
#else
#define SSEL_GLOBAL(NT) static void
#define SSEL_PARAM const
#define SSEL_DEV static inline
#endif

namespace ssel {

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

#ifndef SSEL_HOST
SSEL_DEV void cp_async16(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc) : "memory");
}
SSEL_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
SSEL_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
#endif

// rows 0..R-1 of the tile -> stage[row][32]: thread `tid` copies chunk tid & 7
// (4 nodes) of rows tid >> 3, + NT / 8, ...  A chunk that would start past the
// end of the row (last tile, or a sub-range that is not 32-aligned) re-reads the
// row's last chunk: always inside the allocation, never stored.
template <int NT>
SSEL_DEV void issue_tile(const SselLaunch& L, float* stage, long long tile, int tid) {
    const int ch = tid & 7;
    long long col4 = L.n0 + tile * COLS + ch * 4;
    col4 = col4 > L.pitch - 4 ? L.pitch - 4 : col4;
    const int R = L.a.R;
    int row = tid >> 3;
    const float* src = L.data + static_cast<long long>(row) * L.pitch + col4;
    float* dst = stage + row * COLS + ch * 4;
    const long long sstep = static_cast<long long>(NT / 8) * L.pitch;
    for (; row < R; row += NT / 8) {
        cp_async16(dst, src);
        src += sstep;
        dst += (NT / 8) * COLS;
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
SSEL_GLOBAL(NW * 32)
k2_ssel_kernel(SSEL_PARAM SselLaunch L) {
    constexpr int NT = NW * 32;
    using LO = Layout<NW, IPT>;
    static_assert(LO::BYTES <= 227 * 1024, "shared memory layout too large");
#ifndef SSEL_HOST
    extern __shared__ __align__(128) unsigned char smem_raw[];
#else
    unsigned char* smem_raw = ssel_host_smem(LO::BYTES);
#endif
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
    const int tid = static_cast<int>(threadIdx.x);
    const int w = tid >> 5, c = tid & 31;
    const Args& a = L.a;

    long long tile = static_cast<long long>(blockIdx.x);
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
        const long long next = tile + static_cast<long long>(gridDim.x);
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

}  // namespace ssel
