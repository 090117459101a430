// k2_rsel_kernel.cuh -- the tile loop of the K2 register-tile kernel (see
// k2_grid_rsel.cu for the description and the launcher).  In a header so that
// tests/emul can compile THIS code for the host and run it with one OS thread per
// CUDA thread (tests/emul/k2_rsel_threads.cpp, also under ThreadSanitizer).
// Host build: RSK_HOST is defined by the includer, which supplies the CUDA
// built-ins the kernel uses (tests/emul/cuda_threads_shim.h).
#pragma once
#include "k2_rsel_phases.cuh"

#ifndef RSK_HOST
#define RSK_GLOBAL(NT) __global__ void __launch_bounds__(NT, 1)
#define RSK_PARAM const __grid_constant__
#define RSK_DEV __device__ __forceinline__
#else
#define RSK_GLOBAL(NT) static void
#define RSK_PARAM const
#define RSK_DEV static inline
#endif

namespace rsel {

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

RSK_DEV float ld_stream(const float* p) {
#ifdef __CUDA_ARCH__
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
#else
    return *p;
#endif
}

// rows w, w + NW, ... of column (tile, c): one coalesced 128-byte row segment per
// warp instruction, no branches, nothing that waits for the data (the loads of
// the next tile stay in flight while the current one is finished).  A column
// past the end of the stack re-reads the last column (its planes are never
// stored); a row past R keeps the NaN its register was preset with (only the
// upper half of the row groups can be past R: IPT/2 * NW < R by construction).
RSK_DEV float ld_stream_if(const float* p, bool ok) {
#ifdef __CUDA_ARCH__
    float v = __int_as_float(0x7fc00000);
    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p ld.global.nc.L1::no_allocate.f32 %0, [%1];\n}"
                 : "+f"(v) : "l"(p), "r"(static_cast<unsigned>(ok)));
    return v;
#else
    return ok ? *p : qnan();
#endif
}

template <int NW, int IPT>
RSK_DEV void load_tile(Thr<IPT>& t, const RselLaunch& L, long long tile, int w, int c) {
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
RSK_GLOBAL(NW * 32)
k2_rsel_kernel(RSK_PARAM RselLaunch L) {
    constexpr int NT = NW * 32;
    constexpr int S = NW / 8;
#ifndef RSK_HOST
    extern __shared__ __align__(16) unsigned char smem_raw[];
#else
    unsigned char* smem_raw = gsb_host_smem();
#endif
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

}  // namespace rsel
