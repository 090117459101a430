// k2_csort_kernel.cuh -- the K2 counting-sort kernel itself (see k2_grid_csort.cu for
// the description and the launcher).  It lives in a header so that tests/emul can
// compile THIS code for the host and run it with one OS thread per CUDA thread
// (tests/emul/k2_csort_threads.cpp: warp intrinsics as 32-thread barriers with an
// exchange buffer, the mbarrier and the TMA tile load modelled in software), also
// under ThreadSanitizer -- the closest substitute for compute-sanitizer's racecheck
// on the kernel's hand-rolled ticket / mbarrier / reader-count protocol that the
// GPU pool allows.  Host build: CS_HOST is defined by the includer, which supplies
// the CUDA built-ins the kernel uses.
#pragma once
#include "gsb_common.cuh"
#include "k2_args.cuh"

#ifndef CS_HOST
#define CS_GLOBAL(NT) __global__ void __launch_bounds__(NT, 1)
#define CS_PARAM const __grid_constant__
#define CS_DEV __device__ __forceinline__
#define CS_DEVVAR __device__
#else
#define CS_GLOBAL(NT) static void
#define CS_PARAM const
#define CS_DEV static inline
#define CS_DEVVAR static
#endif

// [0] columns, [1] columns with big bins, [2] columns on the generic path,
// [3] sum of odd-even phases, [4] big bins sorted, [5] cycles waiting for tiles,
// [6] cycles of all warps
CS_DEVVAR unsigned long long gsb_cs_counters[8];

namespace csort_k {

constexpr unsigned FULL = 0xffffffffu;
// warps per CTA: as many as the register file allows for the column depth
template <int IPT> struct CsWarps { static constexpr int v = IPT >= 32 ? 16 : IPT >= 16 ? 24 : 32; };   // 128 / 80 / 64 registers

template <int N> struct CLog2 { static constexpr int v = 1 + CLog2<N / 2>::v; };
template <> struct CLog2<1> { static constexpr int v = 0; };

CS_DEV float cs_warp_sum(float x) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
    return x;
}
CS_DEV float cs_warp_min(float x) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) x = fminf(x, __shfl_xor_sync(FULL, x, o));
    return x;
}
CS_DEV float cs_warp_max(float x) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) x = fmaxf(x, __shfl_xor_sync(FULL, x, o));
    return x;
}

// nodata / NaN / out-of-bounds rows (TMA fills those with NaN) -> +INF.
// setp.ne is the ORDERED not-equal: false for NaN, so one compare covers
// bn.replace(arr, nodata, nan) (grid.py:681) and the NaN dropping.
CS_DEV float cs_sanitize(float x, float nodata) {
#ifdef __CUDA_ARCH__
    float r;
    asm("{\n.reg .pred p;\nsetp.ne.f32 p, %1, %2;\nselp.f32 %0, %1, %3, p;\n}"
        : "=f"(r) : "f"(x), "f"(nodata), "f"(__int_as_float(0x7f800000)));
    return r;
#else
    return (x == x && x != nodata) ? x : __int_as_float(0x7f800000);
#endif
}

CS_DEV uint64_t cs_pack2(float a, float b) {
#ifdef __CUDA_ARCH__
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
#else
    return static_cast<uint64_t>(__float_as_uint(a)) | (static_cast<uint64_t>(__float_as_uint(b)) << 32);
#endif
}
CS_DEV float cs_lo2(uint64_t v) {
    return __uint_as_float(static_cast<uint32_t>(v));
}
CS_DEV float cs_hi2(uint64_t v) {
    return __uint_as_float(static_cast<uint32_t>(v >> 32));
}

// the inline-PTX sites of the column loop, each with the host form the threaded
// emulation (tests/emul/k2_csort_threads.cpp) runs
CS_DEV void cs_min3(float& m, float a, float b) {
#ifdef __CUDA_ARCH__
    asm("min.f32 %0, %0, %1, %2;" : "+f"(m) : "f"(a), "f"(b));
#else
    m = fminf(m, fminf(a, b));
#endif
}
CS_DEV void cs_max3(float& m, float a, float b) {
#ifdef __CUDA_ARCH__
    asm("max.f32 %0, %0, %1, %2;" : "+f"(m) : "f"(a), "f"(b));
#else
    m = fmaxf(m, fmaxf(a, b));
#endif
}
CS_DEV uint64_t cs_sub2(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    uint64_t d;
    asm("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
#else
    return cs_pack2(cs_lo2(a) - cs_lo2(b), cs_hi2(a) - cs_hi2(b));
#endif
}
CS_DEV uint64_t cs_sq2(uint64_t a) {
#ifdef __CUDA_ARCH__
    uint64_t d;
    asm("mul.f32x2 %0, %1, %1;" : "=l"(d) : "l"(a));
    return d;
#else
    return cs_pack2(cs_lo2(a) * cs_lo2(a), cs_hi2(a) * cs_hi2(a));
#endif
}
CS_DEV void cs_acc2(uint64_t& s, uint64_t a) {
#ifdef __CUDA_ARCH__
    asm("add.f32x2 %0, %0, %1;" : "+l"(s) : "l"(a));
#else
    s = cs_pack2(cs_lo2(s) + cs_lo2(a), cs_hi2(s) + cs_hi2(a));
#endif
}
CS_DEV void cs_fma2(uint64_t& s, uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(s) : "l"(a), "l"(b));
#else
    s = cs_pack2(fmaf(cs_lo2(a), cs_lo2(b), cs_lo2(s)), fmaf(cs_hi2(a), cs_hi2(b), cs_hi2(s)));
#endif
}
// shared-memory word at a 32-bit shared address: +1 returning the old value / load
CS_DEV uint32_t cs_atom_inc_shared(uint32_t addr) {
#ifdef __CUDA_ARCH__
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(addr) : "memory");
    return old;
#elif defined(CS_HOST)
    return __atomic_fetch_add(reinterpret_cast<uint32_t*>(gsb_host_shared_ptr(addr)), 1u, __ATOMIC_RELAXED);
#else
    return addr;        // (host pass of nvcc: never called)
#endif
}
CS_DEV uint32_t cs_ld_shared(uint32_t addr) {
#ifdef __CUDA_ARCH__
    uint32_t p;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(p) : "r"(addr) : "memory");
    return p;
#elif defined(CS_HOST)
    return *reinterpret_cast<const uint32_t*>(gsb_host_shared_ptr(addr));
#else
    return addr;        // (host pass of nvcc: never called)
#endif
}

// B(p) = A(p - s) over the NPAD-bit string distributed IPT bits per lane
template <int IPT>
CS_DEV uint32_t cs_bits_shift(uint32_t A, int s, int lane) {
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
CS_GLOBAL(CsWarps<IPT>::v * 32)
k2_csort_kernel(CS_PARAM CUtensorMap tmap, CS_PARAM K2Args a) {
    constexpr int CS_WARPS = CsWarps<IPT>::v;
    constexpr int NPAD = IPT * 32;
    constexpr int LOGI = CLog2<IPT>::v;
    constexpr int LOGN = LOGI + 5;
    // bins: every word of the slot region doubles as a histogram bin, padding
    // words included, so lane l owns the IPT+1 bins [l*(IPT+1), (l+1)*(IPT+1))
    // and bin b simply lives at word b; word B is the bin of the missing values
    constexpr int B = 32 * (IPT + 1);
    // the column maximum and minimum are atoms (DSS clamps to [datamin, datamax],
    // parsers/dss.py:322-325: 65 % of the benchmark columns carry a run of the
    // maximum, mean 66 values, up to 892 of 1000).  Each gets 32 bins, one per
    // lane, so that a warp full of equal values still hits 32 different banks
    // instead of serialising on one address; word INV collects the missing values
    constexpr int HI0 = B, LO0 = B + 32, INV = B + 64;
    constexpr int W = IPT + 2 * HALO;             // window registers per lane
    constexpr int LPAD = K2_CS_LPAD(HALO);
    constexpr int REG_WORDS = K2_CS_REGION_WORDS(IPT, HALO);
    constexpr int TILE_BYTES = NPAD * 128;        // NPAD rows x 32 nodes x 4 B
    constexpr int BOX_ROWS = NPAD < 256 ? NPAD : 256;  // must match make_tmap()
#ifndef CS_HOST
    extern __shared__ __align__(1024) unsigned char smem[];
#else
    unsigned char* smem = gsb_host_smem();
#endif
    uint32_t* scratch_all = reinterpret_cast<uint32_t*>(smem + TILE_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(scratch_all + CS_WARPS * REG_WORDS);
    uint32_t* readers = reinterpret_cast<uint32_t*>(full + 1);
    uint32_t* ticket = readers + 1;
    double* ptab_q = reinterpret_cast<double*>(full + 2);
    int* ptab_kind = reinterpret_cast<int*>(ptab_q + GSB_MAX_PLANES);
    // per warp: [0] number of over-full bins of the current column, [1..96] their
    // (start | count << 16); at most NPAD / (HALO + 2) < 96 of them exist
    uint32_t* biglist = reinterpret_cast<uint32_t*>(ptab_kind + GSB_MAX_PLANES)
                      + (threadIdx.x >> 5) * 100;

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
    auto idx = [](int p) { return p + (p >> LOGI); };

    // tiles: all of [0, ntiles), or the ones listed by the register-tile kernel
    // (k2_grid_rsel.cu) as "could not be finished exactly": [0] count, [1 + k] tile
    const long long nt = a.tile_list ? static_cast<long long>(a.tile_list[0]) : a.ntiles;
    auto phys = [&](long long li) -> long long {
        return a.tile_list ? static_cast<long long>(a.tile_list[1 + li]) : li;
    };
    if (threadIdx.x == 0) {
        *readers = 0u;
        *ticket = 0u;
        gsb_mbar_init(full, 1);
        gsb_fence_barrier_init();
        gsb_prefetch_tmap(&tmap);
        if (blockIdx.x < nt) issue_tile(phys(blockIdx.x));
    }
    // left halo = -INF forever; histogram zero for the first column
    for (int i = lane; i < LPAD; i += 32) A[-1 - i] = -INF;
    for (int i = lane; i < REG_WORDS - LPAD; i += 32) H[i] = 0u;
    __syncthreads();

    // output planes: round r of a column handles plane r*32 + lane; the plane
    // list (kind, moment slot, quantile) is parked in shared memory so that it
    // costs no registers across the column loop
    auto slot_of = [](int kind) {
        return kind == GSB_PK_MEAN ? 1 : kind == GSB_PK_VAR ? 2 : kind == GSB_PK_STD ? 3
             : kind == GSB_PK_COEFVAR ? 4 : kind == GSB_PK_SKEW ? 5 : 1;
    };
    if (threadIdx.x < GSB_MAX_PLANES) {
        const int k = threadIdx.x < a.nplanes ? a.kind[threadIdx.x] : -1;
        ptab_q[threadIdx.x] = threadIdx.x < a.nplanes ? a.q[threadIdx.x] : 0.0;
        ptab_kind[threadIdx.x] = (k < 0) ? -1 : (k | (slot_of(k) << 8));
    }
    __syncthreads();
    const uint32_t hbase = gsb_smem_u32(H);

    // Columns are handed out one at a time from a CTA-wide ticket counter:
    // ticket k is column k % 32 of the CTA's (k / 32)-th tile.  Warps therefore
    // balance themselves (a column with heavy ties costs its warp one ticket,
    // not the whole block a barrier), and the tile buffer goes back to TMA as
    // soon as its 32 columns sit in registers -- about half a tile time before
    // the last of them is finished.
    const long long t_begin = clock64();
    for (;;) {
        {
            uint32_t tk = 0;
            if (lane == 0) tk = atomicAdd(ticket, 1u);
            tk = __shfl_sync(FULL, tk, 0);
            const long long seq = tk >> 5;
            const int c = static_cast<int>(tk & 31u);
            const long long li = blockIdx.x + seq * gridDim.x;
            if (li >= nt) break;
            const long long tile = phys(li);
            if (a.lockstep) {
                const long long t0 = clock64();
                gsb_mbar_wait(full, static_cast<uint32_t>(seq & 1));
                if (lane == 0) atomicAdd(&gsb_cs_counters[5], (unsigned long long)(clock64() - t0));
            } else {
                gsb_mbar_wait(full, static_cast<uint32_t>(seq & 1));
            }
            const long long col = tile * 32 + c;
            // ---- column c of the tile -> registers, nodata/NaN/padding -> +INF
            // (128-byte TMA swizzle: 16-byte chunk index ^ (row & 7))
            const uint32_t my_off = lane * 128u
                + (((static_cast<uint32_t>(c) >> 2) ^ (lane & 7u)) << 4) + (c & 3) * 4u;
            float xs[IPT];
            {
                const unsigned char* tp = smem + my_off;
                if (a.nchunks == IPT) {
#pragma unroll
                    for (int j = 0; j < IPT; ++j)
                        xs[j] = cs_sanitize(*reinterpret_cast<const float*>(tp + j * 4096), a.nodata);
                } else {
#pragma unroll
                    for (int j = 0; j < IPT; ++j)
                        xs[j] = j < a.nchunks
                            ? cs_sanitize(*reinterpret_cast<const float*>(tp + j * 4096), a.nodata) : INF;
                }
            }
            // the warp that loads the tile's last column hands the buffer back
            // to TMA
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                if (atomicAdd(readers, 1u) == 31u) {
                    *readers = 0u;
                    const long long nxt = li + gridDim.x;
                    if (nxt < nt) issue_tile(phys(nxt));
                }
            }

            // ---- pass 0: pivot, shifted moments (packed f32x2), min / max
            // pivot of the shifted sums: an actual element (a constant column gives
            // d == 0 exactly, and the max below may use it as the stand-in for
            // missing values) within a few sigma of the mean, which is all the
            // float32 shifted sums need
            float piv, S1, S2, S3, lo, hi;
            {
                // the element, among the first 32 realizations, nearest to their
                // mean: an element (a constant column gives d == 0 exactly) and
                // within a fraction of sigma of the column mean.  The float32
                // shifted sums lose |pivot - mean| / sigma digits in the third
                // moment: a pivot taken blindly (first / median of three elements)
                // sat 4 sigma out on skewed columns and cost 5e-5 in the skewness.
                // (taken from the first group of 32 realizations that holds at least 8
                // values: with the first 32 missing, the old pivot 0 sat mean / sigma
                // standard deviations out and the skewness of such a column was off
                // by 1e-4 and more -- found by the host emulation and its fuzz,
                // tests/emul/k2_csort_threads.cpp, fuzz_csort.py)
                float xp = xs[0];
                uint32_t okm = __ballot_sync(FULL, xp < INF);
                if (__popc(okm) < 8) {
                    // rare: fewer than 8 of the first 32 realizations exist -> the
                    // first later group of 32 that holds 8, else the fullest one (a
                    // group with one or two values makes a tail value the pivot)
#pragma unroll
                    for (int j = 1; j < IPT; ++j)
                        if (__popc(okm) < 8) {
                            const float xj = xs[j];
                            const uint32_t mj = __ballot_sync(FULL, xj < INF);
                            if (__popc(mj) > __popc(okm)) { okm = mj; xp = xj; }
                        }
                }
                const bool ok0 = xp < INF;
                const bool have_piv = okm != 0u;
                const float m0 = cs_warp_sum(ok0 ? xp : 0.f) / static_cast<float>(__popc(okm));
                const float dist = ok0 ? fabsf(xp - m0) : INF;         // NaN when the sum overflowed
                const uint32_t dkey = (dist == dist) ? __float_as_uint(dist) : 0x7f800000u;
                const uint32_t dmin = __reduce_min_sync(FULL, dkey);
                const uint32_t who = __ballot_sync(FULL, dkey == dmin && ok0);
                const int lsel = who ? __ffs(who) - 1 : (have_piv ? __ffs(okm) - 1 : 0);
                const float cand = __shfl_sync(FULL, xp, lsel);
                piv = have_piv ? cand : 0.f;
                float mn = INF, mx = have_piv ? cand : -INF;
                uint64_t s1 = 0, s2 = 0, s3 = 0;
                const uint64_t pp = cs_pack2(piv, piv);
#pragma unroll
                for (int j = 0; j < IPT; j += 2) {
                    float xm[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) xm[u] = xs[j + u] < INF ? xs[j + u] : piv;
                    // 3-input FMNMX3; piv is an element: a valid stand-in for missing values
                    cs_min3(mn, xs[j], xs[j + 1]);
                    cs_max3(mx, xm[0], xm[1]);
                    const uint64_t xx = cs_pack2(xm[0], xm[1]);
                    const uint64_t d = cs_sub2(xx, pp);
                    const uint64_t d2 = cs_sq2(d);
                    cs_acc2(s1, d);
                    cs_acc2(s2, d2);
                    cs_fma2(s3, d2, d);
                }
                if (!have_piv) {
                    // empty column: the pivot (0) is not an element, take the
                    // maximum the slow way (it stays -INF)
                    float m = -INF;
#pragma unroll
                    for (int j = 0; j < IPT; ++j) m = xs[j] < INF ? fmaxf(m, xs[j]) : m;
                    mx = m;
                }
                S1 = cs_warp_sum(cs_lo2(s1) + cs_hi2(s1));
                S2 = cs_warp_sum(cs_lo2(s2) + cs_hi2(s2));
                S3 = cs_warp_sum(cs_lo2(s3) + cs_hi2(s3));
                lo = cs_warp_min(mn);
                hi = cs_warp_max(mx);
            }
            // ---- pass 1: histogram; the ATOMS result is the arrival index.
            // regular bin = round((x - lo) * scale) + 1 in [1, B-2]; values equal
            // to the column maximum / minimum go to this lane's atom bin (an FMA
            // with a 0/1 factor moves them there); +INF (missing) -> bin INV
            float scale = static_cast<float>(B - 3) / (hi - lo);
            const bool degenerate = !(scale < INF) || !(scale > 0.f);   // range over/underflow
            if (degenerate) scale = 0.f;
            const float lo_eff = degenerate ? 0.f : lo;
            const bool allsame = !(lo < hi);     // constant or empty column: nothing to sort
            const float dhi = static_cast<float>(8388608 + HI0 + lane)
                            - fmaf(hi - lo_eff, scale, 8388609.0f);
            const float dlo = allsame ? 0.f : static_cast<float>(8388608 + LO0 + lane)
                            - fmaf(lo - lo_eff, scale, 8388609.0f);
            uint32_t packed[IPT];
#pragma unroll
            for (int j = 0; j < IPT; ++j) {
                const float x = xs[j];
                float t = fmaf(x - lo_eff, scale, 8388609.0f);
                t = fmaf((x == hi) ? 1.0f : 0.0f, dhi, t);
                t = fmaf((x == lo) ? 1.0f : 0.0f, dlo, t);
                const int b = __viaddmin_s32_relu(__float_as_int(t), -0x4B000000, INV);
                const uint32_t addr = hbase + static_cast<uint32_t>(b) * 4u;
                const uint32_t old = cs_atom_inc_shared(addr);
                packed[j] = old * 262144u + addr;
            }
            __syncwarp();

            // ---- scan: lane l owns bins [l*(IPT+1), (l+1)*(IPT+1))
            // bins of more than HALO+1 values (rare: heavy ties away from the
            // extremes, a dense spot of the distribution) are sorted one by one
            // after the scatter; `maxc` is the largest of the others
            int nc, maxc, n_lo, n_hi;
            int nbig = 0;            // over-full bins in this column
            bool generic = false;
            {
                uint32_t* hp = H + lane * (IPT + 1);
                uint32_t tot = 0, mc = 0;
#pragma unroll
                for (int i0 = 0; i0 <= IPT; i0 += 11) {
                    uint32_t cnt[11];
#pragma unroll
                    for (int u = 0; u < 11; ++u) if (i0 + u <= IPT) cnt[u] = hp[i0 + u];
#pragma unroll
                    for (int u = 0; u < 11; ++u)
                        if (i0 + u <= IPT) { tot += cnt[u]; mc = max(mc, cnt[u]); }
                }
                // exclusive prefixes of the regular bins and of the two atom strips
                const uint32_t c_hi = H[HI0 + lane], c_lo = H[LO0 + lane];
                uint32_t incl = tot, incl_hi = c_hi, incl_lo = c_lo;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(FULL, incl, o);
                    const uint32_t th = __shfl_up_sync(FULL, incl_hi, o);
                    const uint32_t tl = __shfl_up_sync(FULL, incl_lo, o);
                    if (lane >= o) { incl += t; incl_hi += th; incl_lo += tl; }
                }
                n_lo = static_cast<int>(__shfl_sync(FULL, incl_lo, 31));
                n_hi = static_cast<int>(__shfl_sync(FULL, incl_hi, 31));
                const int n_reg = static_cast<int>(__shfl_sync(FULL, incl, 31));
                nc = n_lo + n_reg + n_hi;
                maxc = static_cast<int>(__reduce_max_sync(FULL, mc));
                H[LO0 + lane] = incl_lo - c_lo;
                H[HI0 + lane] = static_cast<uint32_t>(n_lo + n_reg) + incl_hi - c_hi;
                uint32_t run = static_cast<uint32_t>(n_lo) + incl - tot;
                if (maxc <= HALO + 1) {
#pragma unroll
                    for (int i0 = 0; i0 <= IPT; i0 += 11) {
                        uint32_t cnt[11];
#pragma unroll
                        for (int u = 0; u < 11; ++u) if (i0 + u <= IPT) cnt[u] = hp[i0 + u];
#pragma unroll
                        for (int u = 0; u < 11; ++u)
                            if (i0 + u <= IPT) { hp[i0 + u] = run; run += cnt[u]; }
                    }
                } else {
                    // rare: list the over-full bins; `maxc` becomes the largest
                    // of the others
                    if (lane == 0) biglist[0] = 0u;
                    __syncwarp();
                    mc = 0;
#pragma unroll 1
                    for (int i = 0; i <= IPT; ++i) {
                        const uint32_t cn = hp[i];
                        hp[i] = run;
                        if (cn > HALO + 1) {
                            const uint32_t k = atomicAdd(biglist, 1u);
                            if (k < 96u) biglist[1 + k] = run | (cn << 16);
                        } else {
                            mc = max(mc, cn);
                        }
                        run += cn;
                    }
                    maxc = static_cast<int>(__reduce_max_sync(FULL, mc));
                    __syncwarp();
                    nbig = static_cast<int>(biglist[0]);
                }
                if (lane == 0) H[INV] = static_cast<uint32_t>(nc);   // missing values go last
            }
            __syncwarp();

            if (a.lockstep && lane == 0) {     // GSB_K2_STATS=1
                atomicAdd(&gsb_cs_counters[0], 1ull);
                if (nbig) atomicAdd(&gsb_cs_counters[1], 1ull);
                atomicAdd(&gsb_cs_counters[3], (unsigned long long)maxc);
                atomicAdd(&gsb_cs_counters[4], (unsigned long long)nbig);
            }
            // ---- moment finals: warp-uniform, every lane computes them (no
            // shared-memory round trip); the two sums with cancellation in
            // float64, divisions and roots in float32 (1e-7 relative)
            float st_mean, st_var, st_sd, st_cv, st_skew;
            {
                const float fnan = __int_as_float(0x7fc00000);
                const float fn = static_cast<float>(nc);
                const double d1 = static_cast<double>(S1), d2 = static_cast<double>(S2);
                const float delta = __fdiv_rn(S1, fn);
                const double dd = static_cast<double>(delta);
                const bool constant = (S2 == 0.f);
                const float m2 = constant ? 0.f : fmaxf(static_cast<float>(d2 - d1 * dd), 0.f);
                const float m3 = constant ? 0.f : static_cast<float>(
                    static_cast<double>(S3) - 3.0 * dd * d2 + 2.0 * d1 * dd * dd);
                const float mean = piv + delta;
                const float var = __fdiv_rn(m2, fn - 1.f);
                const float sd = sqrtf(var);
                st_mean = nc > 0 ? mean : fnan;
                st_var = nc > 1 ? var : fnan;
                st_sd = nc > 1 ? sd : fnan;
                st_cv = nc > 1 ? __fdiv_rn(sd, mean) * 100.f : fnan;
                st_skew = nc < 3 ? fnan : (m2 == 0.f) ? 0.f
                        : __fdiv_rn(fn * sqrtf(fn - 1.f), fn - 2.f) * __fdiv_rn(m3, m2 * sqrtf(m2));
            }
            float w[W];
            if (!allsame) {
                // ---- pass 2: slot = prefix[bin] + arrival, then scatter
#pragma unroll
                for (int j = 0; j < IPT; ++j) {
                    const uint32_t p = cs_ld_shared(packed[j] & 0x3ffffu);
                    packed[j] = p + (packed[j] >> 18);
                }
                __syncwarp();
                if (lane < HALO) A[idx(NPAD + lane)] = INF;          // right halo (aliases the atom strips)
#pragma unroll
                for (int j = 0; j < IPT; ++j) A[idx(static_cast<int>(packed[j]))] = xs[j];
                __syncwarp();

                // ---- over-full bins, the warp on one bin at a time.  A bin whose
                // values are all equal (heavy ties: quantised data) is already in
                // order; a mixed one is rank-sorted; a mixed one of more than 128
                // values (an outlier collapsed the bin map) sends the column to
                // the generic sort
#pragma unroll 1
                for (int e = 0; e < nbig && !generic; ++e) {
                    const uint32_t ent = biglist[1 + e];
                    const int s0 = static_cast<int>(ent & 0xffffu);
                    const int cn = static_cast<int>(ent >> 16);
                    float bmn = INF, bmx = -INF;
#pragma unroll 1
                    for (int i = lane; i < cn; i += 32) {
                        const float y = A[idx(s0 + i)];
                        bmn = fminf(bmn, y);
                        bmx = fmaxf(bmx, y);
                    }
                    bmn = cs_warp_min(bmn);
                    bmx = cs_warp_max(bmx);
                    if (bmn == bmx) continue;
                    if (cn > 128) { generic = true; break; }
                    float mine[4];
                    int rank[4];
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const int i = lane + 32 * m;
                        mine[m] = i < cn ? A[idx(s0 + i)] : INF;
                        rank[m] = 0;
                    }
#pragma unroll 1
                    for (int jx = 0; jx < cn; ++jx) {
                        const float y = A[idx(s0 + jx)];
#pragma unroll
                        for (int m = 0; m < 4; ++m)
                            rank[m] += (y < mine[m] || (y == mine[m] && jx < lane + 32 * m)) ? 1 : 0;
                    }
                    __syncwarp();
#pragma unroll
                    for (int m = 0; m < 4; ++m)
                        if (lane + 32 * m < cn) A[idx(s0 + rank[m])] = mine[m];
                    __syncwarp();
                }
                if (a.lockstep && generic && lane == 0) atomicAdd(&gsb_cs_counters[2], 1ull);
                if (generic) {
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
                    for (int ph = maxc > 1 ? maxc : 0; ph > 0; ph -= 2) {
#pragma unroll
                        for (int j = 0; j + 1 < W; j += 2) {
                            const float x = w[j], y = w[j + 1];
                            w[j] = fminf(x, y);
                            w[j + 1] = fmaxf(x, y);
                        }
                        if (ph == 1) break;
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
                    int pe = 0, best = 1;
                    if (generic || nbig) {
                        // runs of any length: binary lifting
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
                        if (top >= 0) {
                            const uint32_t bl = __ballot_sync(FULL, cur != 0);
                            const int lf = __ffs(bl) - 1;
                            const uint32_t wd = __shfl_sync(FULL, cur, lf);
                            pe = lf * IPT + __ffs(wd) - 1;
                        }
                        modev = A[idx(pe)];
                    } else {
                        // window path: the extremes sit in their own bins, so
                        // their multiplicities are the bin counts; every other
                        // run is at most maxc long -- grow it one step at a time.
                        // Keep only bits of the regular positions (n_lo, nc - n_hi)
                        {
                            const int a0 = n_lo + 1 - lane * IPT;      // first kept bit
                            const int a1 = nc - n_hi - lane * IPT;     // first dropped bit
                            uint32_t keep = 0xffffffffu;
                            if (a0 >= 32) keep = 0; else if (a0 > 0) keep &= ~((1u << a0) - 1u);
                            if (a1 <= 0) keep = 0; else if (a1 < 32) keep &= (1u << a1) - 1u;
                            M &= keep;
                        }
                        uint32_t cur = M;      // bit p: a run of >= 2 ends at p
                        int k = 1;             // run length - 1 of the bits in `cur`
                        if (__any_sync(FULL, cur != 0)) {
#pragma unroll 1
                            for (;;) {
                                const uint32_t nx = cur & cs_bits_shift<IPT>(M, k, lane);
                                if (!__any_sync(FULL, nx != 0)) break;
                                cur = nx;
                                ++k;
                            }
                            const uint32_t bl = __ballot_sync(FULL, cur != 0);
                            const int lf = __ffs(bl) - 1;
                            const uint32_t wd = __shfl_sync(FULL, cur, lf);
                            pe = lf * IPT + __ffs(wd) - 1 - k;     // start of the first longest run
                            best = k + 1;
                        } else {
                            pe = n_lo < nc ? n_lo : 0;             // all regular values distinct
                        }
                        // ties -> smallest value: lo atom, then regular, then hi atom
                        modev = A[idx(pe)];
                        if (n_lo >= best) { modev = lo; best = n_lo; }
                        if (n_hi > best) modev = hi;
                    }
                }
            }

            // every lane runs the quantile arithmetic (no divergence between the
            // quantile, moment and mode lanes) and selects its plane at the end
            auto plane_value = [&](int kind, int slot, double q) -> float {
                const float fnan = __int_as_float(0x7fc00000);
                // np.quantile(method='linear'): h = (n-1) q, NumPy _lerp
                const double h = __dmul_rn(dn - 1.0, q);
                const double fl = floor(h);
                int l0 = static_cast<int>(fl);
                l0 = max(0, min(l0, nc - 2));
                const double t = h - fl;
                float f0 = vmin, f1 = vmin;
                if (!allsame && nc >= 2) { f0 = A[idx(l0)]; f1 = A[idx(l0 + 1)]; }
                const double x0 = static_cast<double>(f0), x1 = static_cast<double>(f1);
                const double d = x1 - x0;
                const double rq = (t >= 0.5) ? __dsub_rn(x1, __dmul_rn(d, 1.0 - t))
                                             : __dadd_rn(x0, __dmul_rn(d, t));
                float r = static_cast<float>(rq);
                if (h >= dn - 1.0) r = vmax;
                if (h < 0.0 || allsame) r = vmin;
                float m = slot == 1 ? st_mean : slot == 2 ? st_var : slot == 3 ? st_sd
                        : slot == 4 ? st_cv : st_skew;
                if (allsame && slot >= 2) {
                    if (slot == 4) m = (nc > 1) ? 0.f / st_mean * 100.f : fnan;
                    else if (slot == 5) m = (nc >= 3) ? 0.f : fnan;
                    else m = (nc > 1) ? 0.f : fnan;
                }
                r = kind == GSB_PK_QUANT ? r : kind == GSB_PK_MODE ? modev : m;
                return nc == 0 ? fnan : r;
            };
            if (col >= a.skip && col < a.nn) {
#pragma unroll 1
                for (int pl = lane; pl < a.nplanes; pl += 32) {
                    const int kk = ptab_kind[pl];
                    a.out[static_cast<long long>(pl) * a.out_pitch + col] =
                        plane_value(kk & 0xff, kk >> 8, ptab_q[pl]);
                }
            }
            __syncwarp();
            // histogram back to zero for the next column
            {
                float4* z = reinterpret_cast<float4*>(H);
                for (int i = lane; i < (INV + 4) / 4; i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncwarp();
        }
    }
    if (a.lockstep && lane == 0)
        atomicAdd(&gsb_cs_counters[6], (unsigned long long)(clock64() - t_begin));
}

}  // namespace csort_k
