// k2_grid_csort.cu -- K2, counting-sort variant: realization-axis reduction
// with exact order statistics for 64 < R <= 1024.
//
// Same contract as k2_grid_sort.cu (GridFiles.stats cell loop,
// GSIMCLI/tools/grid.py:662-701): per grid node, R realization values ->
// moments, any number of linearly interpolated quantiles, and the mode.
//
// A comparison sort of a 1024-value column costs ~110 instructions per value
// (k2_grid_sort.cu: 11 % of the HBM roofline).  This kernel sorts by
// DISTRIBUTION instead:
//   HBM stack [R][pitch] f32 --TMA 32-node boxes, 128B swizzle--> one 128 KB
//   tile per CTA --LDS.64, conflict free--> registers (warp w owns nodes
//   2w, 2w+1; lane l holds rows 32 j + l), tile buffer refilled by TMA at once.
//   pass 0  registers: nodata/NaN -> +INF, shifted moments, column min/max
//   pass 1  bin = linear map of [min, max] onto B = NPAD bins; one shared-memory
//           ATOMS.ADD per value returns its arrival index inside the bin
//   scan    per-bin exclusive prefix (lane l owns bins l*IPT..; stride IPT+1
//           padding keeps every LDS/STS bank-conflict free)
//   pass 2  value -> slot prefix[bin] + arrival (the slot array aliases the
//           histogram: prefixes are read into registers first)
//   windows lane l re-reads slots [l*IPT - HALO, (l+1)*IPT + HALO) into
//           registers and runs max-bin-size odd-even transposition phases:
//           bins are value-ordered, so exchanges never cross a bin boundary
//           and every bin <= HALO+1 values that touches the lane's own slots
//           is sorted completely inside the window
//   columns whose largest bin exceeds HALO+1 (heavy ties, outliers) take a
//   generic shared-memory bitonic sort instead -- always exact.
//   finish  quantile lerp in float64 (NumPy _lerp), mode from the equal-
//           neighbour bitmap, moments in float64; lane = output plane.
// HBM traffic: the stack is read exactly once (4 B per node-realization).
#include "gsb_common.cuh"
#include "k2_args.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int CS_WARPS = 16;

template <int N> struct CLog2 { static constexpr int v = 1 + CLog2<N / 2>::v; };
template <> struct CLog2<1> { static constexpr int v = 0; };

__device__ __forceinline__ float cs_warp_sum(float x) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
    return x;
}
__device__ __forceinline__ float cs_warp_min(float x) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) x = fminf(x, __shfl_xor_sync(FULL, x, o));
    return x;
}
__device__ __forceinline__ float cs_warp_max(float x) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) x = fmaxf(x, __shfl_xor_sync(FULL, x, o));
    return x;
}

// nodata / NaN / out-of-bounds rows (TMA fills those with NaN) -> +INF.
// setp.ne is the ORDERED not-equal: false for NaN, so one compare covers
// bn.replace(arr, nodata, nan) (grid.py:681) and the NaN dropping.
__device__ __forceinline__ float cs_sanitize(float x, float nodata) {
    float r;
    asm("{\n.reg .pred p;\nsetp.ne.f32 p, %1, %2;\nselp.f32 %0, %1, %3, p;\n}"
        : "=f"(r) : "f"(x), "f"(nodata), "f"(__int_as_float(0x7f800000)));
    return r;
}

// B(p) = A(p - s) over the NPAD-bit string distributed IPT bits per lane
template <int IPT>
__device__ __forceinline__ uint32_t cs_bits_shift(uint32_t A, int s, int lane) {
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

template <int IPT, bool MODE, int HALO>
__global__ void __launch_bounds__(CS_WARPS * 32, 1)
k2_csort_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ K2Args a) {
    constexpr int NPAD = IPT * 32;
    constexpr int LOGI = CLog2<IPT>::v;
    constexpr int LOGN = LOGI + 5;
    constexpr int B = NPAD;                       // bins (+1 overflow bin for missing values)
    constexpr int W = IPT + 2 * HALO;             // window registers per lane
    constexpr int LPAD = K2_CS_LPAD(HALO);
    constexpr int REG_WORDS = K2_CS_REGION_WORDS(IPT, HALO);
    constexpr int TILE_BYTES = NPAD * 128;        // NPAD rows x 32 nodes x 4 B
    constexpr int BOX_ROWS = NPAD < 256 ? NPAD : 256;  // must match make_tmap()
    extern __shared__ __align__(1024) unsigned char smem[];
    uint32_t* scratch_all = reinterpret_cast<uint32_t*>(smem + TILE_BYTES);
    float* cstat_all = reinterpret_cast<float*>(scratch_all + CS_WARPS * REG_WORDS);
    uint64_t* full = reinterpret_cast<uint64_t*>(cstat_all + CS_WARPS * 8);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nbox = (a.R + BOX_ROWS - 1) / BOX_ROWS;
    const float INF = __int_as_float(0x7f800000);

    auto issue_tile = [&](long long tile) {
        const int col = static_cast<int>(a.n0 + tile * 32);
        gsb_fence_proxy_async();
        gsb_mbar_expect_tx(full, static_cast<uint32_t>(nbox) * BOX_ROWS * 128u);
        for (int b = 0; b < nbox; ++b)
            gsb_tma_load_2d(smem + b * BOX_ROWS * 128, &tmap, full, col, b * BOX_ROWS);
    };

    // slot array A and histogram H share the warp's region; position p lives at
    // word p + (p >> LOGI): one padding word per IPT slots, so that lanes that
    // own consecutive IPT-slot groups sit IPT+1 words apart (odd => all 32
    // banks distinct)
    float* A = reinterpret_cast<float*>(scratch_all + warp * REG_WORDS + LPAD);
    uint32_t* H = reinterpret_cast<uint32_t*>(A);
    float* cstat = cstat_all + warp * 8;
    auto idx = [](int p) { return p + (p >> LOGI); };

    if (threadIdx.x == 0) {
        gsb_mbar_init(full, 1);
        gsb_fence_barrier_init();
        gsb_prefetch_tmap(&tmap);
        if (blockIdx.x < a.ntiles) issue_tile(blockIdx.x);
    }
    // left halo = -INF forever; histogram zero for the first column
    for (int i = lane; i < LPAD; i += 32) A[-1 - i] = -INF;
    for (int i = lane; i < REG_WORDS - LPAD; i += 32) H[i] = 0u;
    __syncthreads();

    // bytes of this lane's 8 bytes (nodes 2w, 2w+1 of row `lane`) inside a
    // 32-row chunk, 128-byte TMA swizzle applied
    const uint32_t my_off = lane * 128u
        + (((static_cast<uint32_t>(warp) >> 1) ^ (lane & 7u)) << 4) + (warp & 1u) * 8u;

    // this lane's output planes (round r handles plane r*32 + lane)
    int kind0 = -1, kind1 = -1;
    double q0 = 0.0, q1 = 0.0;
    if (lane < a.nplanes) { kind0 = a.kind[lane]; q0 = a.q[lane]; }
    if (lane + 32 < a.nplanes) { kind1 = a.kind[lane + 32]; q1 = a.q[lane + 32]; }
    auto slot_of = [](int kind) {
        return kind == GSB_PK_MEAN ? 1 : kind == GSB_PK_VAR ? 2 : kind == GSB_PK_STD ? 3
             : kind == GSB_PK_COEFVAR ? 4 : kind == GSB_PK_SKEW ? 5 : 1;
    };
    const int slot0 = slot_of(kind0), slot1 = slot_of(kind1);
    const uint32_t hbase = gsb_smem_u32(H);

    uint32_t phase = 0;
    for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        gsb_mbar_wait(full, phase);
        phase ^= 1u;

        // ---- tile -> registers, nodata/NaN/padding -> +INF
        float xs[IPT], xs1[IPT];
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            if (j < a.nchunks) {
                const float2 t = *reinterpret_cast<const float2*>(smem + j * 4096 + my_off);
                xs[j] = cs_sanitize(t.x, a.nodata);
                xs1[j] = cs_sanitize(t.y, a.nodata);
            } else {
                xs[j] = INF;
                xs1[j] = INF;
            }
        }
        // tile consumed: stream the next one in behind the compute
        __syncthreads();
        if (threadIdx.x == 0) {
            const long long nxt = tile + gridDim.x;
            if (nxt < a.ntiles) issue_tile(nxt);
        }

        // ---- pass 0: pivot, shifted moments, min / max of both columns
        float piv[2], S1[2], S2[2], S3[2], mn[2], mx[2];
        int have_piv[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            // pivot: the valid value among the first 32 realizations closest to
            // their mean (an actual element: the max below may then use it
            // as the stand-in for missing values)
            const float x0 = c == 0 ? xs[0] : xs1[0];
            const bool ok = x0 < INF;
            const float s0 = cs_warp_sum(ok ? x0 : 0.f);
            const int c0 = __reduce_add_sync(FULL, ok ? 1 : 0);
            const float m0 = c0 ? s0 / static_cast<float>(c0) : 0.f;
            float dist = ok ? fabsf(x0 - m0) : INF;
            float cand = ok ? x0 : 0.f;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                const float od = __shfl_xor_sync(FULL, dist, o);
                const float oc = __shfl_xor_sync(FULL, cand, o);
                if (od < dist || (od == dist && oc < cand)) { dist = od; cand = oc; }
            }
            piv[c] = c0 ? cand : 0.f;
            have_piv[c] = c0;
            S1[c] = S2[c] = S3[c] = 0.f;
            mn[c] = INF;
            mx[c] = c0 ? cand : -INF;
        }
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const float x = c == 0 ? xs[j] : xs1[j];
                const bool ok = x < INF;
                const float xm = ok ? x : piv[c];
                mn[c] = fminf(mn[c], x);
                mx[c] = fmaxf(mx[c], xm);        // piv is an element: a valid stand-in
                const float d = xm - piv[c];
                const float d2 = d * d;
                S1[c] += d;
                S2[c] += d2;
                S3[c] = fmaf(d2, d, S3[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            if (!have_piv[c]) {
                // the first 32 realizations are all missing: the pivot (0) is not
                // an element, take the maximum the slow way
                float m = -INF;
#pragma unroll
                for (int j = 0; j < IPT; ++j) {
                    const float x = c == 0 ? xs[j] : xs1[j];
                    m = x < INF ? fmaxf(m, x) : m;
                }
                mx[c] = m;
            }
            S1[c] = cs_warp_sum(S1[c]);
            S2[c] = cs_warp_sum(S2[c]);
            S3[c] = cs_warp_sum(S3[c]);
            mn[c] = cs_warp_min(mn[c]);
            mx[c] = cs_warp_max(mx[c]);
        }

        const long long colbase = tile * 32 + warp * 2;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            if (c == 1) {
#pragma unroll
                for (int j = 0; j < IPT; ++j) xs[j] = xs1[j];
                piv[0] = piv[1]; S1[0] = S1[1]; S2[0] = S2[1]; S3[0] = S3[1];
                mn[0] = mn[1]; mx[0] = mx[1];
            }
            const float lo = mn[0], hi = mx[0];
            // ---- pass 1: histogram; the ATOMS result is the arrival index.
            // bin = round((x - lo) * scale) + 1 in [1, B-3]; the column extremes
            // get bins of their own (0 and B-2): DSS clamps to [datamin, datamax]
            // (parsers/dss.py:322-325), so they are atoms -- often hundreds of
            // equal values -- that need no sorting; +INF (missing) -> bin B
            float scale = static_cast<float>(B - 4) / (hi - lo);
            const bool degenerate = !(scale < INF) || !(scale > 0.f);   // range over/underflow
            if (degenerate) scale = 0.f;
            const float lo_eff = degenerate ? 0.f : lo;
            uint32_t packed[IPT];
#pragma unroll
            for (int j = 0; j < IPT; ++j) {
                const float x = xs[j];
                float t = fmaf(x - lo_eff, scale, 8388609.0f);
                t += (x == hi) ? 1.0f : 0.0f;
                t -= (x == lo) ? 1.0f : 0.0f;
                const int b = __viaddmin_s32_relu(__float_as_int(t), -0x4B000000, B);
                const uint32_t addr = hbase + static_cast<uint32_t>(idx(b)) * 4u;
                uint32_t old;
                asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(addr) : "memory");
                packed[j] = (old << 18) | addr;
            }
            __syncwarp();

            // ---- scan: lane l owns bins [l*IPT, (l+1)*IPT)
            int nc, maxc;
            {
                uint32_t tot = 0, mc = 0;
                volatile uint32_t* hp = H + lane * (IPT + 1);
#pragma unroll
                for (int i = 0; i < IPT; ++i) {
                    const uint32_t cn = hp[i];
                    tot += cn;
                    // the two atom bins hold equal values only: nothing to sort
                    const bool atom = (i == 0 && lane == 0) || (i == IPT - 2 && lane == 31);
                    mc = max(mc, atom ? 0u : cn);
                }
                uint32_t incl = tot;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += t;
                }
                nc = static_cast<int>(__shfl_sync(FULL, incl, 31));
                maxc = static_cast<int>(__reduce_max_sync(FULL, mc));
                uint32_t run = incl - tot;
#pragma unroll
                for (int i = 0; i < IPT; ++i) { const uint32_t cn = hp[i]; hp[i] = run; run += cn; }
                if (lane == 0) H[idx(B)] = static_cast<uint32_t>(nc);   // missing values go last
            }
            __syncwarp();

            // ---- moment finals in float64 (lane 0), parked for the plane lanes
            if (lane == 0) {
                const float fnan = __int_as_float(0x7fc00000);
                const double dn = static_cast<double>(nc);
                const double d1 = static_cast<double>(S1[0]), d2 = static_cast<double>(S2[0]);
                const double delta = d1 / dn;
                const double mean = static_cast<double>(piv[0]) + delta;
                const bool constant = (S2[0] == 0.f);
                const double m2 = constant ? 0.0 : fmax(d2 - d1 * delta, 0.0);
                const double m3 = constant ? 0.0
                    : static_cast<double>(S3[0]) - 3.0 * delta * d2 + 2.0 * dn * delta * delta * delta;
                const double var = m2 / (dn - 1.0);
                const double sd = sqrt(var);
                cstat[0] = __int_as_float(nc);
                cstat[1] = nc > 0 ? static_cast<float>(mean) : fnan;
                cstat[2] = nc > 1 ? static_cast<float>(var) : fnan;
                cstat[3] = nc > 1 ? static_cast<float>(sd) : fnan;
                cstat[4] = nc > 1 ? static_cast<float>(sd / mean * 100.0) : fnan;
                cstat[5] = nc < 3 ? fnan
                         : (m2 == 0.0) ? 0.f
                         : static_cast<float>((dn * sqrt(dn - 1.0) / (dn - 2.0)) * (m3 / (m2 * sqrt(m2))));
            }

            __syncwarp();
            const bool allsame = !(lo < hi);     // constant or empty column: nothing to sort
            float w[W];
            if (!allsame) {
                // ---- pass 2: slot = prefix[bin] + arrival, then scatter
#pragma unroll
                for (int j = 0; j < IPT; ++j) {
                    uint32_t p;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(p) : "r"(packed[j] & 0x3ffffu) : "memory");
                    packed[j] = p + (packed[j] >> 18);
                }
                __syncwarp();
                if (lane < HALO) A[idx(NPAD + lane)] = INF;          // right halo
#pragma unroll
                for (int j = 0; j < IPT; ++j) A[idx(static_cast<int>(packed[j]))] = xs[j];
                __syncwarp();

                if (maxc > HALO + 1) {
                    // ---- generic path: bitonic sort of the slot array in shared memory
#pragma unroll 1
                    for (int k = 2; k <= NPAD; k <<= 1)
#pragma unroll 1
                        for (int jj = k >> 1; jj > 0; jj >>= 1) {
#pragma unroll 1
                            for (int t = lane; t < NPAD / 2; t += 32) {
                                const int i = ((t & ~(jj - 1)) << 1) | (t & (jj - 1));
                                const int p = i | jj;
                                const float x = A[idx(i)], y = A[idx(p)];
                                if ((x > y) == ((i & k) == 0)) { A[idx(i)] = y; A[idx(p)] = x; }
                            }
                            __syncwarp();
                        }
#pragma unroll
                    for (int j = 0; j < IPT; ++j) w[HALO + j] = A[lane * (IPT + 1) + j];
                } else {
                    // ---- windows: [l*IPT - HALO, (l+1)*IPT + HALO) -> registers
                    const float* wp = A + lane * (IPT + 1);
#pragma unroll
                    for (int j = 0; j < W; ++j) {
                        const int r = j - HALO;                       // slot relative to l*IPT
                        const int off = r + (r >= 0 ? r / IPT : -((-r + IPT - 1) / IPT));
                        w[j] = wp[off];
                    }
                    // max-bin-size phases of odd-even transposition
#pragma unroll 1
                    for (int ph = 0; ph < maxc && maxc > 1; ph += 2) {
#pragma unroll
                        for (int j = 0; j + 1 < W; j += 2) {
                            const float x = w[j], y = w[j + 1];
                            w[j] = fminf(x, y);
                            w[j + 1] = fmaxf(x, y);
                        }
#pragma unroll
                        for (int j = 1; j + 1 < W; j += 2) {
                            const float x = w[j], y = w[j + 1];
                            w[j] = fminf(x, y);
                            w[j + 1] = fmaxf(x, y);
                        }
                    }
                    __syncwarp();
                    float* wq = A + lane * (IPT + 1);
#pragma unroll
                    for (int j = 0; j < IPT; ++j) wq[j] = w[HALO + j];
                }
                __syncwarp();
            }

            // ---- planes
            const double dn = static_cast<double>(nc);
            const float vmin = lo;
            const float vmax = allsame ? lo : A[idx(nc > 0 ? nc - 1 : 0)];
            float modev = vmin;
            if constexpr (MODE) {
                if (!allsame) {
                    // bit j: sorted[p] == sorted[p-1], p = lane*IPT + j, p < nc
                    uint32_t M = 0;
                    const float prev_last = __shfl_up_sync(FULL, w[HALO + IPT - 1], 1);
                    if (lane > 0 && w[HALO] == prev_last) M |= 1u;
#pragma unroll
                    for (int j = 1; j < IPT; ++j)
                        if (w[HALO + j] == w[HALO + j - 1]) M |= (1u << j);
                    const int mloc = nc - lane * IPT;
                    if (mloc <= 0) M = 0;
                    else if (mloc < IPT) M &= (1u << mloc) - 1u;
                    uint32_t P[LOGN + 1];
                    P[0] = M;
                    int top = __any_sync(FULL, M != 0) ? 0 : -1;
#pragma unroll
                    for (int i = 0; i < LOGN; ++i) {
                        P[i + 1] = 0;
                        if (top == i) {
                            const uint32_t nx = P[i] & cs_bits_shift<IPT>(P[i], 1 << i, lane);
                            if (__any_sync(FULL, nx != 0)) { P[i + 1] = nx; top = i + 1; }
                        }
                    }
                    uint32_t cur = 0;
                    int len = 0;
#pragma unroll
                    for (int i = LOGN; i >= 0; --i) {
                        if (i == top) { cur = P[i]; len = 1 << i; }
                        else if (i < top) {
                            const uint32_t cand = cur & cs_bits_shift<IPT>(P[i], len, lane);
                            if (__any_sync(FULL, cand != 0)) { cur = cand; len += 1 << i; }
                        }
                    }
                    int pe = 0;
                    if (top >= 0) {
                        const uint32_t bl = __ballot_sync(FULL, cur != 0);
                        const int lf = __ffs(bl) - 1;
                        const uint32_t wd = __shfl_sync(FULL, cur, lf);
                        pe = lf * IPT + __ffs(wd) - 1;
                    }
                    modev = A[idx(pe)];
                }
            }

            auto plane_value = [&](int kind, int slot, double q) -> float {
                const float fnan = __int_as_float(0x7fc00000);
                if (nc == 0) return fnan;
                if (kind == GSB_PK_QUANT) {
                    if (allsame) return vmin;
                    // np.quantile(method='linear'): h = (n-1) q, NumPy _lerp
                    const double h = __dmul_rn(dn - 1.0, q);
                    if (h >= dn - 1.0) return vmax;
                    if (h < 0.0) return vmin;
                    const double fl = floor(h);
                    const int l0 = static_cast<int>(fl);
                    const double t = h - fl;
                    const double x0 = static_cast<double>(A[idx(l0)]);
                    const double x1 = static_cast<double>(A[idx(l0 + 1)]);
                    const double d = x1 - x0;
                    const double r = (t >= 0.5) ? __dsub_rn(x1, __dmul_rn(d, 1.0 - t))
                                                : __dadd_rn(x0, __dmul_rn(d, t));
                    return static_cast<float>(r);
                }
                if (kind == GSB_PK_MODE) return modev;
                float r = cstat[slot];
                if (allsame && slot >= 2) {
                    if (slot == 4) r = (nc > 1) ? 0.f / cstat[1] * 100.f : fnan;
                    else if (slot == 5) r = (nc >= 3) ? 0.f : fnan;
                    else r = (nc > 1) ? 0.f : fnan;
                }
                return r;
            };
            const long long col = colbase + c;
            if (col >= a.skip && col < a.nn) {
                if (kind0 >= 0)
                    a.out[static_cast<long long>(lane) * a.out_pitch + col] =
                        plane_value(kind0, slot0, q0);
                if (kind1 >= 0)
                    a.out[static_cast<long long>(lane + 32) * a.out_pitch + col] =
                        plane_value(kind1, slot1, q1);
            }
            __syncwarp();
            // histogram back to zero for the next column
            {
                float4* z = reinterpret_cast<float4*>(H);
                for (int i = lane; i < (NPAD + 32 + 4) / 4; i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncwarp();
        }
    }
}

template <int IPT, bool MODE, int HALO>
int cs_launch_one(gsb_stack* s, const K2Args& args, int grid, cudaStream_t st) {
    constexpr int NPAD = IPT * 32;
    const size_t smem = static_cast<size_t>(NPAD) * 128
        + static_cast<size_t>(CS_WARPS) * K2_CS_REGION_WORDS(IPT, HALO) * 4 + CS_WARPS * 8 * 4 + 16;
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
    case 8: return cs_launch_one<8, MODE, HALO>(s, args, grid, st);
    case 16: return cs_launch_one<16, MODE, HALO>(s, args, grid, st);
    case 32: return cs_launch_one<32, MODE, HALO>(s, args, grid, st);
    }
    gsb_set_error("k2 csort: unsupported items-per-lane %d", ipt);
    return GSB_EINVAL;
}

}  // namespace

int gsb_launch_k2_csort(gsb_stack* s, const K2Args& args, int ipt, int grid, bool mode,
                        int halo, cudaStream_t st) {
    if (halo == 12)
        return mode ? cs_launch_ipt<true, 12>(s, args, ipt, grid, st)
                    : cs_launch_ipt<false, 12>(s, args, ipt, grid, st);
    if (halo == 8)
        return mode ? cs_launch_ipt<true, 8>(s, args, ipt, grid, st)
                    : cs_launch_ipt<false, 8>(s, args, ipt, grid, st);
    return mode ? cs_launch_ipt<true, 10>(s, args, ipt, grid, st)
                : cs_launch_ipt<false, 10>(s, args, ipt, grid, st);
}
