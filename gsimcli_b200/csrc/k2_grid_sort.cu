// k2_grid_sort.cu -- K2: realization-axis reduction with exact order statistics
//
// Replaces the cell loop of GridFiles.stats (GSIMCLI/tools/grid.py:662-701):
// for every grid node, the R realization values are reduced to mean / variance
// / std / coefvar / skewness (moments) and to any number of linearly
// interpolated quantiles (median, percentile pair, percentile lists) and the
// mode (extension).
//
// Data flow (one persistent CTA of 8 warps per SM; 256 threads so that each
// thread may use up to 255 registers -- the column tile lives in registers):
//   HBM stack [R][pitch] f32  --TMA 2-D boxes 32 rows x 32 nodes, 128B swizzle-->
//   32-stage (128 KB) shared-memory ring, refilled by whichever warp is the
//   last to finish reading a stage  --LDS.128, conflict-free-->
//   registers: warp w owns nodes 4w..4w+3 of the tile, lane l holds rows
//   {32 j + l}; IPT = ceil32(R)/32 values per lane and node.
// In registers: (1) nodata/NaN -> +INF and shifted one-pass moments (pivot =
// mean of the first 32 realizations, so no cancellation); (2) a bitonic sort
// of the padded column across the warp: Batcher odd-even merge sort inside
// each lane, then bitonic merge phases whose cross-lane stages are one
// SHFL + one FMNMX per element thanks to a per-lane sign state (a lane that
// must keep the larger value stores negated values, so every lane executes
// y = min(y, -y_partner)); (3) the sorted column is parked in shared memory,
// each lane computes one output plane (quantile lerp in float64 exactly as
// NumPy's _lerp; moments finished in float64; mode from a run-length bitmap
// with binary lifting) and stores 4 adjacent nodes with one 16-byte store.
// HBM traffic: the stack is read exactly once (4 B per node-realization).
#include "gsb_common.cuh"
#include "k2_args.cuh"
#include <utility>
#include <stdlib.h>

namespace {

constexpr int K2_STAGE_BYTES = 4096;   // 32 rows x 32 nodes x 4 B
constexpr int K2_NST = 32;             // ring depth: one full R=1024 tile
constexpr unsigned FULL = 0xffffffffu;


// ---------------------------------------------------------------------------
// Batcher odd-even merge sort network for N = 2^k inputs, generated at compile
// time (191 compare-exchanges for N = 32; bitonic would need 240).
template <int N>
struct OemNet {
    static constexpr int MAXCE = (N <= 1) ? 1 : N * 8;
    int n = 0;
    int a[MAXCE] = {};
    int b[MAXCE] = {};
};

template <int N>
constexpr OemNet<N> make_oem() {
    OemNet<N> net{};
    for (int p = 1; p < N; p <<= 1)
        for (int k = p; k >= 1; k >>= 1)
            for (int j = k % p; j + k < N; j += 2 * k)
                for (int i = 0; i < k && i + j + k < N; ++i)
                    if ((i + j) / (2 * p) == (i + j + k) / (2 * p)) {
                        net.a[net.n] = i + j;
                        net.b[net.n] = i + j + k;
                        ++net.n;
                    }
    return net;
}

template <int IPT>
struct Oem {
    static constexpr OemNet<IPT> net = make_oem<IPT>();
};

template <int I, int J, int CPW, int IPT>
__device__ __forceinline__ void ce4(float (&v)[CPW][IPT]) {
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
        const float x = v[c][I], y = v[c][J];
        v[c][I] = fminf(x, y);
        v[c][J] = fmaxf(x, y);
    }
}

// All 8 warps execute the same straight-line sorting code (~250 KB of SASS
// per tile at IPT = 32).  Left alone they drift apart and each streams the
// code through the instruction caches on its own: ncu showed `no_instruction`
// as the top stall (2.9 per issue).  A block barrier every few hundred
// instructions keeps the warps inside one I-cache window, so a line is
// fetched once per SM instead of once per warp.
__device__ __forceinline__ void lockstep() { __syncthreads(); }

template <int I>
__device__ __forceinline__ void lockstep_every(bool on) {
    if constexpr ((I % 48) == 47) { if (on) lockstep(); }
}

template <int CPW, int IPT, int... Is>
__device__ __forceinline__ void oem_apply(float (&v)[CPW][IPT], bool ls,
                                          std::integer_sequence<int, Is...>) {
    ((ce4<Oem<IPT>::net.a[Is], Oem<IPT>::net.b[Is], CPW, IPT>(v), lockstep_every<Is>(ls)), ...);
}

template <int CPW, int IPT>
__device__ __forceinline__ void sort_in_lane(float (&v)[CPW][IPT], bool ls) {
    if constexpr (IPT > 1)
        oem_apply<CPW, IPT>(v, ls, std::make_integer_sequence<int, Oem<IPT>::net.n>{});
}

// v *= f with f = +-1.  Inline PTX keeps it an FMUL (fma pipe, idle here):
// NVVM would rewrite v * select(-1, 1) into FSEL(v, -v) on the alu pipe, which
// is the pipe the FMNMX compare-exchanges saturate.
__device__ __forceinline__ float mul_sign(float x, float f) {
    float r;
    asm("mul.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(f));
    return r;
}

template <int CPW, int IPT>
__device__ __forceinline__ void scale_all(float (&v)[CPW][IPT], float f) {
#pragma unroll
    for (int c = 0; c < CPW; ++c)
#pragma unroll
        for (int j = 0; j < IPT; ++j) v[c][j] = mul_sign(v[c][j], f);
}



// Bitonic merge phases across the warp.  Logical position = lane*IPT + reg.
// On entry every lane is sorted ascending in ITS stored values, and lanes with
// (lane & 1) hold negated values (i.e. are descending in x).  `neg` tracks the
// lane's sign state.  On exit all lanes hold plain x, globally ascending.
template <int CPW, int IPT>
__device__ __forceinline__ void merge_across_lanes(float (&v)[CPW][IPT], int lane,
                                                   bool neg, bool ls_on) {
    // the phase and stage loops stay ROLLED: one copy of the cross-lane stage
    // body (IPT*4 x {FMUL, SHFL, FMNMX}) and one of the in-lane half-cleaner
    // block, re-executed from the instruction cache 15 and 5 times per tile
#pragma unroll 1
    for (int L = 2; L <= 32; L <<= 1) {
        const bool desc = (L < 32) && ((lane & L) != 0);
#pragma unroll 1
        for (int ls = L >> 1; ls >= 1; ls >>= 1) {
            // the lane that keeps the larger value works on negated data
            const bool want = ((lane & ls) != 0) != desc;
            scale_all<CPW, IPT>(v, (want != neg) ? -1.0f : 1.0f);
            neg = want;
#pragma unroll
            for (int c = 0; c < CPW; ++c)
#pragma unroll
                for (int j = 0; j < IPT; ++j) {
                    const float o = __shfl_xor_sync(FULL, v[c][j], ls);
                    v[c][j] = fminf(v[c][j], -o);
                }
            if constexpr (IPT >= 8) { if (ls_on) lockstep(); }
        }
        if constexpr (IPT > 1) {
            scale_all<CPW, IPT>(v, (desc != neg) ? -1.0f : 1.0f);
            neg = desc;
#pragma unroll
            for (int s = IPT >> 1; s >= 1; s >>= 1)
#pragma unroll
                for (int j = 0; j < IPT; ++j)
                    if ((j & s) == 0) {
#pragma unroll
                        for (int c = 0; c < CPW; ++c) {
                            const float x = v[c][j], y = v[c][j | s];
                            v[c][j] = fminf(x, y);
                            v[c][j | s] = fmaxf(x, y);
                        }
                    }
            if constexpr (IPT >= 8) { if (ls_on) lockstep(); }
        } else {
            scale_all<CPW, IPT>(v, (desc != neg) ? -1.0f : 1.0f);
            neg = desc;
        }
    }
}

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
    return x;
}

// B(p) = A(p - s) over the NPAD-bit string distributed IPT bits per lane
template <int IPT>
__device__ __forceinline__ uint32_t bits_shift(uint32_t A, int s, int lane) {
    constexpr uint32_t WM = (IPT == 32) ? 0xffffffffu : ((1u << IPT) - 1u);
    const int ws = s / IPT, bs = s % IPT;
    uint32_t lo = __shfl_up_sync(FULL, A, ws & 31);
    uint32_t hi = __shfl_up_sync(FULL, A, (ws + 1) & 31);
    if (lane < ws || ws > 31) lo = 0;
    if (lane < ws + 1 || ws + 1 > 31) hi = 0;
    uint32_t r = lo << bs;
    if (bs) r |= hi >> (IPT - bs);
    return r & WM;
}

// scratch address of sorted position p (rotation keeps the dump conflict-free)
template <int IPT>
__device__ __forceinline__ int scr_addr(int p) {
    const int l = p / IPT, j = p % IPT;
    return l * IPT + ((j + ((l * IPT) >> 5)) & (IPT - 1));
}

// nodata / NaN / out-of-bounds (TMA fills those with NaN) -> +INF, everything
// carrying the lane's initial sign.  setp.ne is the ORDERED not-equal: false
// for NaN, so one compare covers bn.replace(arr, nodata, nan) (grid.py:681)
// and NaN dropping.
__device__ __forceinline__ float sanitize(float x, float nodata, float sg, float infs) {
    float r;
    const float xs = mul_sign(x, sg);
    asm("{\n.reg .pred p;\nsetp.ne.f32 p, %1, %2;\nselp.f32 %0, %3, %4, p;\n}"
        : "=f"(r) : "f"(x), "f"(nodata), "f"(xs), "f"(infs));
    return r;
}
__device__ __forceinline__ bool is_valid(float x, float nodata) {
    int r;
    asm("{\n.reg .pred p;\nsetp.ne.f32 p, %1, %2;\nselp.s32 %0, 1, 0, p;\n}"
        : "=r"(r) : "f"(x), "f"(nodata));
    return r != 0;
}

template <int IPT>
__device__ __forceinline__ void dump_column(float* scr, const float (&col)[IPT], int lane) {
#pragma unroll
    for (int j = 0; j < IPT; ++j)
        scr[lane * IPT + ((j + ((lane * IPT) >> 5)) & (IPT - 1))] = col[j];
}

// bit j of lane l: sorted[p] == sorted[p-1], p = l*IPT + j
template <int IPT>
__device__ __forceinline__ uint32_t equal_mask(const float (&col)[IPT], int lane) {
    uint32_t M = 0;
    const float prev_last = __shfl_up_sync(FULL, col[IPT - 1], 1);
    if (lane > 0 && col[0] == prev_last) M |= 1u;
#pragma unroll
    for (int j = 1; j < IPT; ++j)
        if (col[j] == col[j - 1]) M |= (1u << j);
    return M;
}

// CPW = nodes (columns) per warp: 4 -> 8 warps x 255 registers, 2 -> 16 warps x
// 128 registers (more warps per scheduler to cover SHFL / FMNMX latencies).
template <int IPT, bool MODE, int CPW>
__global__ void __launch_bounds__(32 / CPW * 32, 1)
k2_sort_kernel(const __grid_constant__ CUtensorMap tmap,
               const __grid_constant__ K2Args a) {
    constexpr int K2_WARPS = 32 / CPW;
    constexpr int NPAD = IPT * 32;
    constexpr int TILE_BYTES = NPAD * 128;            // NPAD rows x 32 nodes x 4 B
    constexpr int BOX_ROWS = NPAD < 256 ? NPAD : 256; // must match make_tmap()
    constexpr int LOGN = (NPAD == 32) ? 5 : (NPAD == 64) ? 6 : (NPAD == 128) ? 7
                       : (NPAD == 256) ? 8 : (NPAD == 512) ? 9 : 10;
    extern __shared__ __align__(1024) unsigned char smem[];
    float* scratch_all = reinterpret_cast<float*>(smem + TILE_BYTES);
    float* cstat_all = scratch_all + K2_WARPS * NPAD;
    uint64_t* full = reinterpret_cast<uint64_t*>(cstat_all + K2_WARPS * 32);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nbox = (a.R + BOX_ROWS - 1) / BOX_ROWS;

    // one TMA transaction group per tile: nbox boxes of BOX_ROWS x 32 nodes,
    // all completing on the same mbarrier
    auto issue_tile = [&](long long tile) {
        const int col = static_cast<int>(a.n0 + tile * 32);
        gsb_fence_proxy_async();
        gsb_mbar_expect_tx(full, static_cast<uint32_t>(nbox) * BOX_ROWS * 128u);
        for (int b = 0; b < nbox; ++b)
            gsb_tma_load_2d(smem + b * BOX_ROWS * 128, &tmap, full, col, b * BOX_ROWS);
    };

    if (threadIdx.x == 0) {
        gsb_mbar_init(full, 1);
        gsb_fence_barrier_init();
        gsb_prefetch_tmap(&tmap);
        if (blockIdx.x < a.ntiles) issue_tile(blockIdx.x);
    }
    __syncthreads();

    float* scr = scratch_all + warp * NPAD;
    float* cstat = cstat_all + warp * 32;
    // byte offset of this lane's 16 bytes (nodes 4w..4w+3 of row `lane`) inside
    // a 32-row chunk, with the 128-byte TMA swizzle applied
    const uint32_t my_off = (CPW == 4)
        ? lane * 128u + ((static_cast<uint32_t>(warp) ^ (lane & 7u)) << 4)
        : lane * 128u + (((static_cast<uint32_t>(warp) >> 1) ^ (lane & 7u)) << 4) + (warp & 1u) * 8u;
    auto load_row = [&](int j, float (&xs)[CPW]) {
        const unsigned char* ptr = smem + j * K2_STAGE_BYTES + my_off;
        if constexpr (CPW == 4) {
            const float4 t = *reinterpret_cast<const float4*>(ptr);
            xs[0] = t.x; xs[1] = t.y; xs[2] = t.z; xs[3] = t.w;
        } else {
            const float2 t = *reinterpret_cast<const float2*>(ptr);
            xs[0] = t.x; xs[1] = t.y;
        }
    };
    const float INF = __int_as_float(0x7f800000);
    const float sg = (lane & 1) ? -1.0f : 1.0f;
    const float infs = (lane & 1) ? -INF : INF;

    // this lane's output planes (round r handles plane r*32 + lane)
    int kind0 = -1, kind1 = -1;
    double q0 = 0.0, q1 = 0.0;
    if (lane < a.nplanes) { kind0 = a.kind[lane]; q0 = a.q[lane]; }
    if (lane + 32 < a.nplanes) { kind1 = a.kind[lane + 32]; q1 = a.q[lane + 32]; }
    // where the parked moment results of a column live: cstat[c*8 + slot]
    auto slot_of = [](int kind) {
        return kind == GSB_PK_MEAN ? 1 : kind == GSB_PK_VAR ? 2 : kind == GSB_PK_STD ? 3
             : kind == GSB_PK_COEFVAR ? 4 : kind == GSB_PK_SKEW ? 5 : 1;
    };
    const int slot0 = slot_of(kind0), slot1 = slot_of(kind1);

    uint32_t phase = 0;
    for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        gsb_mbar_wait(full, phase);
        phase ^= 1u;

        // ---- (1) tile -> registers, nodata/NaN/padding -> +INF (straight-line)
        float v[CPW][IPT];
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            if (j < a.nchunks) {
                float xs[CPW];
                load_row(j, xs);
#pragma unroll
                for (int c = 0; c < CPW; ++c) v[c][j] = sanitize(xs[c], a.nodata, sg, infs);
            } else {
#pragma unroll
                for (int c = 0; c < CPW; ++c) v[c][j] = infs;
            }
        }

        // ---- (2) shifted one-pass moments straight from the tile (rolled loop)
        float piv[CPW], S1[CPW], S2[CPW], S3[CPW];
        int cnt[CPW];
        {
            float xs[CPW];
            load_row(0, xs);
#pragma unroll
            for (int c = 0; c < CPW; ++c) {
                // pivot: the valid value among the first 32 realizations that is
                // closest to their mean.  Close to the column mean (no
                // cancellation in the shifted sums) and an actual element, so
                // a constant column gives d == 0 and m2 == m3 == 0 exactly.
                const bool ok = is_valid(xs[c], a.nodata);
                const float s0 = warp_sum(ok ? xs[c] : 0.f);
                const int c0 = __reduce_add_sync(FULL, ok ? 1 : 0);
                const float m0 = c0 ? s0 / static_cast<float>(c0) : 0.f;
                float dist = ok ? fabsf(xs[c] - m0) : INF;
                float cand = ok ? xs[c] : 0.f;
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) {
                    const float od = __shfl_xor_sync(FULL, dist, o);
                    const float oc = __shfl_xor_sync(FULL, cand, o);
                    if (od < dist || (od == dist && oc < cand)) { dist = od; cand = oc; }
                }
                piv[c] = c0 ? cand : 0.f;
                S1[c] = S2[c] = S3[c] = 0.f;
                cnt[c] = 0;
            }
        }
#pragma unroll 1
        for (int j = 0; j < a.nchunks; ++j) {
            float xs[CPW];
            load_row(j, xs);
#pragma unroll
            for (int c = 0; c < CPW; ++c) {
                const bool ok = is_valid(xs[c], a.nodata);
                const float d = ok ? xs[c] - piv[c] : 0.f;
                const float d2 = d * d;
                cnt[c] += ok ? 1 : 0;
                S1[c] += d;
                S2[c] += d2;
                S3[c] = fmaf(d2, d, S3[c]);
            }
        }

        // ---- in-lane sort + release of the tile.  Warps 4..7 sort their lanes
        // BEFORE the block barrier, warps 0..3 after it, so the two warps that
        // share a scheduler (w and w+4) run about one merge phase apart: one is
        // in a SHFL-bound cross-lane stage while the other is in an FMNMX-bound
        // in-lane stage.  (One copy of the network: the loop is not unrolled.)
#pragma unroll 1
        for (int rep = 0; rep < 2; ++rep) {
            if (((warp & 4) != 0) == (rep == 0) || a.skew == 0) {
                if (a.skew != 0 || rep == 1) sort_in_lane<CPW, IPT>(v, a.lockstep != 0);
            }
            if (rep == 0) {
                // tile consumed by every warp: stream the next one in behind the sort
                __syncthreads();
                if (threadIdx.x == 0) {
                    const long long nxt = tile + gridDim.x;
                    if (nxt < a.ntiles) issue_tile(nxt);
                }
            }
        }

        // ---- moment finals, once per tile: lane c finishes column c in float64
        // and parks the five float32 results the plane lanes will pick up
        {
            int n_l = 0;
            float p_l = 0.f, s1_l = 0.f, s2_l = 0.f, s3_l = 0.f;
#pragma unroll
            for (int c = 0; c < CPW; ++c) {
                const int n = __reduce_add_sync(FULL, cnt[c]);
                const float s1 = warp_sum(S1[c]);
                const float s2 = warp_sum(S2[c]);
                const float s3 = warp_sum(S3[c]);
                if (lane == c) { n_l = n; p_l = piv[c]; s1_l = s1; s2_l = s2; s3_l = s3; }
            }
            if (lane < CPW) {
                const float fnan = __int_as_float(0x7fc00000);
                const double dn = static_cast<double>(n_l);
                const double d1 = static_cast<double>(s1_l), d2 = static_cast<double>(s2_l);
                const double delta = d1 / dn;
                const double mean = static_cast<double>(p_l) + delta;
                const bool constant = (s2_l == 0.f);
                const double m2 = constant ? 0.0 : fmax(d2 - d1 * delta, 0.0);
                const double m3 = constant ? 0.0
                    : static_cast<double>(s3_l) - 3.0 * delta * d2 + 2.0 * dn * delta * delta * delta;
                const double var = m2 / (dn - 1.0);
                const double sd = sqrt(var);
                float* o = cstat + lane * 8;
                o[0] = __int_as_float(n_l);
                o[1] = n_l > 0 ? static_cast<float>(mean) : fnan;
                o[2] = n_l > 1 ? static_cast<float>(var) : fnan;
                o[3] = n_l > 1 ? static_cast<float>(sd) : fnan;
                o[4] = n_l > 1 ? static_cast<float>(sd / mean * 100.0) : fnan;
                o[5] = n_l < 3 ? fnan
                     : (m2 == 0.0) ? 0.f
                     : static_cast<float>((dn * sqrt(dn - 1.0) / (dn - 2.0)) * (m3 / (m2 * sqrt(m2))));
            }
        }

        // ---- (3) merge the sorted lanes (lane-major positions, ascending)
        merge_across_lanes<CPW, IPT>(v, lane, (lane & 1) != 0, a.lockstep != 0);

        uint32_t M0 = 0, M1 = 0, M2 = 0, M3 = 0;
        if constexpr (MODE) {
            M0 = equal_mask<IPT>(v[0], lane);
            M1 = equal_mask<IPT>(v[1], lane);
            if constexpr (CPW == 4) {
                M2 = equal_mask<IPT>(v[2], lane);
                M3 = equal_mask<IPT>(v[3], lane);
            }
        }

        // ---- (4) per column (rolled): park sorted values, evaluate the planes
        const long long colbase = tile * 32 + warp * CPW;
#pragma unroll 1
        for (int c = 0; c < CPW; ++c) {
            __syncwarp();
            if constexpr (CPW == 4) {
                switch (c) {
                case 0: dump_column<IPT>(scr, v[0], lane); break;
                case 1: dump_column<IPT>(scr, v[1], lane); break;
                case 2: dump_column<IPT>(scr, v[2], lane); break;
                default: dump_column<IPT>(scr, v[CPW - 1], lane); break;
                }
            } else {
                if (c == 0) dump_column<IPT>(scr, v[0], lane);
                else dump_column<IPT>(scr, v[CPW - 1], lane);
            }
            __syncwarp();

            const int nc = __float_as_int(cstat[c * 8 + 0]);
            const double dn = static_cast<double>(nc);
            const float vmin = scr[scr_addr<IPT>(0)];
            const float vmax = scr[scr_addr<IPT>(nc > 0 ? nc - 1 : 0)];
            // a column whose first 32 realizations are all missing has pivot 0:
            // enforce the exact zeros of a constant column from the sorted ends
            const bool constant = (vmin == vmax);

            float modev = 0.f;
            if constexpr (MODE) {
                uint32_t M = (c == 0) ? M0 : (c == 1) ? M1 : (c == 2) ? M2 : M3;
                const int mloc = nc - lane * IPT;     // valid positions in this lane
                if (mloc <= 0) M = 0;
                else if (mloc < IPT) M &= (1u << mloc) - 1u;
                uint32_t P[LOGN + 1];
                P[0] = M;
                int top = __any_sync(FULL, M != 0) ? 0 : -1;
#pragma unroll
                for (int i = 0; i < LOGN; ++i) {
                    P[i + 1] = 0;
                    if (top == i) {
                        const uint32_t nx = P[i] & bits_shift<IPT>(P[i], 1 << i, lane);
                        if (__any_sync(FULL, nx != 0)) { P[i + 1] = nx; top = i + 1; }
                    }
                }
                uint32_t cur = 0;
                int len = 0;
#pragma unroll
                for (int i = LOGN; i >= 0; --i) {
                    if (i == top) { cur = P[i]; len = 1 << i; }
                    else if (i < top) {
                        const uint32_t cand = cur & bits_shift<IPT>(P[i], len, lane);
                        if (__any_sync(FULL, cand != 0)) { cur = cand; len += 1 << i; }
                    }
                }
                int pe = 0;
                if (top >= 0) {
                    const uint32_t b = __ballot_sync(FULL, cur != 0);
                    const int lf = __ffs(b) - 1;
                    const uint32_t w = __shfl_sync(FULL, cur, lf);
                    pe = lf * IPT + __ffs(w) - 1;
                }
                modev = scr[scr_addr<IPT>(pe)];
            }

            auto plane_value = [&](int kind, int slot, double q) -> float {
                const float fnan = __int_as_float(0x7fc00000);
                if (nc == 0) return fnan;
                if (kind == GSB_PK_QUANT) {
                    // np.quantile(method='linear'): h = (n-1) q, NumPy _lerp
                    const double h = __dmul_rn(dn - 1.0, q);
                    if (h >= dn - 1.0) return vmax;
                    if (h < 0.0) return vmin;
                    const double fl = floor(h);
                    const int lo = static_cast<int>(fl);
                    const double t = h - fl;
                    const double x0 = static_cast<double>(scr[scr_addr<IPT>(lo)]);
                    const double x1 = static_cast<double>(scr[scr_addr<IPT>(lo + 1)]);
                    const double d = x1 - x0;
                    const double r = (t >= 0.5) ? __dsub_rn(x1, __dmul_rn(d, 1.0 - t))
                                                : __dadd_rn(x0, __dmul_rn(d, t));
                    return static_cast<float>(r);
                }
                if (kind == GSB_PK_MODE) return modev;
                float r = cstat[c * 8 + slot];
                if (constant && slot >= 2) {
                    if (slot == 4) r = (nc > 1) ? 0.f / cstat[c * 8 + 1] * 100.f : fnan;
                    else if (slot == 5) r = (nc >= 3) ? 0.f : fnan;
                    else r = (nc > 1) ? 0.f : fnan;
                }
                return r;
            };
            // lane = plane: neighbouring warps/columns complete the sectors in L2
            const long long col = colbase + c;
            if (col >= a.skip && col < a.nn) {
                if (kind0 >= 0)
                    a.out[static_cast<long long>(lane) * a.out_pitch + col] =
                        plane_value(kind0, slot0, q0);
                if (kind1 >= 0)
                    a.out[static_cast<long long>(lane + 32) * a.out_pitch + col] =
                        plane_value(kind1, slot1, q1);
            }
        }
    }
}

template <int IPT, bool MODE, int CPW>
int launch_one(gsb_stack* s, const K2Args& args, int grid, cudaStream_t st) {
    constexpr int NPAD = IPT * 32;
    constexpr int WARPS = 32 / CPW;
    const size_t smem = static_cast<size_t>(NPAD) * 128 + WARPS * NPAD * 4 + WARPS * 32 * 4 + 16;
    auto kern = k2_sort_kernel<IPT, MODE, CPW>;
    GSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
    kern<<<grid, WARPS * 32, smem, st>>>(s->tmap_k2, args);
    GSB_CUDA(cudaGetLastError());
    return GSB_OK;
}

template <bool MODE, int CPW>
int launch_ipt(gsb_stack* s, const K2Args& args, int ipt, int grid, cudaStream_t st) {
    switch (ipt) {
    case 1: return launch_one<1, MODE, CPW>(s, args, grid, st);
    case 2: return launch_one<2, MODE, CPW>(s, args, grid, st);
    case 4: return launch_one<4, MODE, CPW>(s, args, grid, st);
    case 8: return launch_one<8, MODE, CPW>(s, args, grid, st);
    case 16: return launch_one<16, MODE, CPW>(s, args, grid, st);
    case 32: return launch_one<32, MODE, CPW>(s, args, grid, st);
    }
    gsb_set_error("k2: unsupported items-per-lane %d", ipt);
    return GSB_EINVAL;
}

}  // namespace

int gsb_k2_box_rows(int64_t R) {
    int npad = 32;
    while (npad < R) npad <<= 1;
    return npad < 256 ? npad : 256;
}

int gsb_launch_k2_sort(gsb_stack* s, const GsbPlanes& pl, float nodata, int64_t n0,
                       int64_t nn, float* d_out, int64_t out_pitch, cudaStream_t st) {
    GSB_REQUIRE(s->R <= GSB_MAX_R_GRID, "k2 sort kernel supports R <= %d (got %lld)",
                GSB_MAX_R_GRID, (long long)s->R);
    GSB_REQUIRE(s->tmap_ok, "stack has no tensor map (driver entry point missing)");
    K2Args args;
    memset(&args, 0, sizeof(args));
    args.R = static_cast<int>(s->R);
    args.nchunks = static_cast<int>((s->R + 31) / 32);
    args.nplanes = pl.n;
    args.nodata = nodata;
    args.lockstep = s->k2_stats;      // diagnostic counters of the counting-sort kernel
    args.skew = 1;
    const int64_t head = n0 & 3;
    args.n0 = n0 - head;
    args.nn = nn + head;
    args.skip = head;
    args.ntiles = (args.nn + 31) / 32;
    args.out = d_out - head;
    args.out_pitch = out_pitch;
    for (int i = 0; i < pl.n; ++i) { args.kind[i] = pl.kind[i]; args.q[i] = pl.q[i]; }
    int ipt = 1;
    while (ipt * 32 < s->R) ipt <<= 1;
    const int ctas = s->sm_count - s->k2_sm_reserve;
    const int grid = static_cast<int>(args.ntiles < ctas ? args.ntiles : ctas);
    if (grid <= 0) return GSB_OK;
    // R > 32: the counting-sort kernel (k2_grid_csort.cu) -- except the
    // reference's own call shape (grid.py:601-602: moments + median + one
    // percentile pair, i.e. at most 3 quantile planes and no mode) on deep
    // columns, where the register-tile kernel (k2_grid_rsel.cu, 16 warps) is
    // 10-20 % faster (R = 200 / 500 / 1000: 4.70 / 3.06 / 3.18 ms against 5.30 /
    // 3.90 / 3.69 ms per 4 GB of stack, profiles/r02_k2_stock_call_timing.log);
    // with more quantile planes or the mode it is not.
    // R <= 32: the register bitonic sort below.  `k2_algo`
    // (gsb_stack_set_option, tests and profiling only) forces one of the three
    // implementations so that the parity suite can cross-check them: 1 =
    // counting sort, 2 = bitonic for every R, 3 = register tile.  Tiles the
    // register-tile kernel flags (bin map collapsed by an outlier, heavy ties,
    // range overflow) are redone by the counting-sort kernel on the same stream.
    int nquant = 0;
    for (int i = 0; i < pl.n; ++i) nquant += pl.kind[i] == GSB_PK_QUANT ? 1 : 0;
    const bool stock_call = !pl.need_mode && nquant <= 3 && s->R >= 200;
    const int algo = s->k2_algo;
    if (ipt >= 2 && algo != 2) {
        if (algo == 1 || (algo == 0 && !stock_call))
            return gsb_launch_k2_csort(s, args, ipt, grid, pl.need_mode != 0, 10, st);
        const int nw = algo == 3 ? (s->k2_nw == 16 ? 16 : 32) : 16;
        int* d_flagged = s->k2_flags;        // ntiles <= pitch / 32 + 1 (unaligned n0)
        GSB_CUDA(cudaMemsetAsync(d_flagged, 0, sizeof(int), st));
        // 4 = the staged-select kernel (k2_grid_ssel.cu): experimental, never the
        // default; shapes it does not serve fall to the register-tile kernel
        const bool ssel = algo == 4 && gsb_k2_ssel_supports(args, pl.need_mode != 0);
        int rc = ssel ? gsb_launch_k2_ssel(s, args, grid, false, s->k2_nw == 32 ? 32 : 16, d_flagged, st)
                      : gsb_launch_k2_rsel(s, args, grid, pl.need_mode != 0, nw, d_flagged, st);
        if (rc) return rc;
        args.tile_list = d_flagged;
        return gsb_launch_k2_csort(s, args, ipt, grid, pl.need_mode != 0, 10, st);
    }
    return pl.need_mode ? launch_ipt<true, 2>(s, args, ipt, grid, st)
                        : launch_ipt<false, 2>(s, args, ipt, grid, st);
}
