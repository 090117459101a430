// k2_ssel_phases.cuh -- the per-thread phases of the K2 "staged select" kernel
// (k2_grid_ssel.cu): GridFiles.stats cell loop, GSIMCLI/tools/grid.py:662-701,
// for the reference's own call shape (moments + at most 3 quantile planes, no
// mode, grid.py:601-602) on deep columns.  EXPERIMENTAL: written after the
// round's GPU budget was spent, so it is opt-in (stack option k2_algo = 4) and
// never chosen by the default dispatch; its algorithm is checked on the CPU by
// tests/emul (this header compiled by g++), and bench.py times it in a
// subprocess and compares it bit for bit with the default kernel.
//
// Same tile and ownership as the register-tile kernel (k2_rsel_phases.cuh):
// 32 adjacent nodes x R realizations per CTA, LANE = COLUMN, thread (warp w,
// lane c) holds rows w, w + NW, ... of column c, every shared-memory structure
// is [index][32 columns] (conflict free).  What changes is where the next tile
// waits and what happens after the histogram:
//   * the next tile is copied global -> shared memory by cp.async (no registers,
//     no instruction after the issue) while the current one, already in
//     registers, is reduced: the whole tile time hides the fetch, where the
//     register-tile kernel has only its (short) finish phase;
//   * select, don't sort: with <= 3 quantiles at most 6 ranks are wanted.  The
//     bins that hold them are found from the raw histogram (two-level walk, no
//     scan, the bin is claimed for a list by compare-and-swap), and one more
//     pass over the registers copies only the values of claimed bins into small
//     per-column lists ([list][slot][column]); the scatter of every value and the
//     125 KB sort buffer are gone, which is what frees the room for the staged
//     tile.  A fixed sorting network on the list gives the order statistics.
// Phases between CTA barriers:
//   A      pivot (same rule as the register-tile kernel), stage -> registers,
//          nodata/NaN -> NaN, shifted power sums, min/max (atomics on ordered keys)
//   [the copy of the next tile is issued here: the stage has been read]
//   C1+B   range -> bin map; float64 sums of the partials; histogram (RED)
//   S1     totals of NW chunks of bins; list counters cleared
//   L      one warp per wanted rank: chunk walk, bin walk, claim the bin
//   X      values of claimed bins -> lists (clamp atoms are counted by difference)
//   O      planes -> global memory; histogram cleared
// Exactness: as in the register-tile kernel, bins are value-ordered and every
// comparison is on the original float32.  A list that overflows (a wanted bin
// with more than CAP regular values) or a range that overflows float32 flags the
// tile, which the counting-sort kernel then redoes -- never a guess.
#pragma once
#include "k2_rsel_phases.cuh"

namespace ssel {

using rsel::Args;
using rsel::COLS;
using rsel::NBINS;
using rsel::NCELLS;
using rsel::TRASH;
using rsel::bin_of;
using rsel::dadd;
using rsel::dmul;
using rsel::dsub;
using rsel::f2u;
using rsel::okey;
using rsel::okey_inv;
using rsel::pinf;
using rsel::qnan;
using rsel::u2f;
using rsel::valid;
using rsel::PK_COEFVAR;
using rsel::PK_MEAN;
using rsel::PK_QUANT;
using rsel::PK_SKEW;
using rsel::PK_STD;
using rsel::PK_VAR;

constexpr int CAP = 32;                       // values a list can hold
constexpr int LSTRIDE = (CAP + 1) * COLS;     // words per list: [count][CAP values] x 32 columns
constexpr int NLMAX = 6;                      // lists = wanted ranks = 2 x quantile planes
constexpr uint32_t MARK_SHIFT = 16;           // cells[b] = count | (list + 1) << 16
constexpr uint32_t COUNT_MASK = 0xffffu;

// rank record ([NLMAX][32]): where one wanted rank lives
//   bits 0..10 offset inside the bin, 11..21 population of the bin, 22..25 list,
//   26..27 kind (0 inner bin, 1 first bin, 2 last bin), 31 valid
constexpr uint32_t RI_VALID = 0x80000000u;

struct Smem {
    float* stage;        // [NW * IPT][32]  the tile as it arrives (rows >= R stay NaN)
    uint32_t* cells;     // [NCELLS][32]    histogram (+ list marks in the high half)
    uint32_t* part;      // [3][NW][32]     float32 partial power sums (aliases `lists`)
    uint32_t* lists;     // [NLMAX][CAP + 1][32]
    uint32_t* rinfo;     // [NLMAX][32]
    uint32_t* chunk;     // [NW][32]        bin totals per chunk of CH bins
    uint32_t* mnmx;      // [2][32]         ordered keys of the column minimum / maximum
    double* msum;        // [3][32]         float64 power sums
    uint32_t* ncol;      // [32]            valid values of the column
};

template <int IPT>
struct Thr {
    float xs[IPT];
    float piv, lo, hi, scale;
    int dead;            // range overflow: the tile is flagged
};

RS_HD void atomic_min_u32(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
    atomicMin(p, v);
#else
    uint32_t o = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
#endif
}
RS_HD uint32_t atomic_cas_u32(uint32_t* p, uint32_t expect, uint32_t v) {
#ifdef __CUDA_ARCH__
    return atomicCAS(p, expect, v);
#else
    uint32_t o = expect;
    __atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
    return o;                      // the old value: `expect` on success
#endif
}

// a word other warps may be claiming by compare-and-swap at the same time (phase L:
// only the mark in the high half changes, the reader masks it off)
RS_HD uint32_t ld_racy(const uint32_t* p) {
#ifdef __CUDA_ARCH__
    return *reinterpret_cast<const volatile uint32_t*>(p);
#else
    return __atomic_load_n(p, __ATOMIC_RELAXED);
#endif
}

// first valid value among the rows of warp k (rows k, k + NW, ...) of column c,
// NaN when there is none: what phase_pivot_write of the register-tile kernel
// stores for that warp
template <int NW, int IPT>
RS_HD float first_valid_of_warp(const float* col, int k, float nodata) {
    float v = col[k * COLS];
    if (valid(v, nodata)) return v;
    for (int i = 1; i < IPT; ++i) {
        v = col[(k + NW * i) * COLS];
        if (valid(v, nodata)) return v;
    }
    return qnan();
}

// ------------------------------------------------------------------ phase A
// Pivot = the lower median of the candidates (first valid value of warps 0..7)
// -- rsel::pivot_of_candidates, read straight from the stage so
// that no exchange and no barrier is needed; every thread of the column
// computes the same value.  Then the thread's rows go to registers and into the
// shifted sums; the pivot is an element of the column, so it stands in for
// missing values in the sums (d == 0), the minimum and the maximum.
template <int NW, int IPT>
RS_HD void phase_a(Thr<IPT>& t, const Smem& s, const Args& a, int w, int c) {
    static_assert(IPT % 2 == 0, "IPT must be even");
    const float* col = s.stage + c;
    // rsel::pivot_of_candidates (lower median of the first 8 valid per-warp
    // candidates), the candidates read straight from the stage
    const float nodata = a.nodata;
    const float p = rsel::pivot_of_candidates<NW>(
        [col, nodata](int k) { return first_valid_of_warp<NW, IPT>(col, k, nodata); });
    t.piv = p;
    const float piv = p;
    float mn = piv, mx = piv;
    const float* src = col + w * COLS;
#ifdef __CUDA_ARCH__
    unsigned long long s1 = 0, s2 = 0, s3 = 0, pp;
    asm("mov.b64 %0, {%1, %1};" : "=l"(pp) : "f"(piv));
#pragma unroll
    for (int i = 0; i < IPT; i += 2) {
        const float r0 = src[i * NW * COLS], r1 = src[(i + 1) * NW * COLS];
        // valid(r) = r is not NaN and r != nodata = the ORDERED not-equal compare
        // (a.nodata is never NaN: the launcher passes +INF for a NaN nodata), taken
        // as an all-ones / zero mask so that both selects are plain bit logic: no
        // predicate to keep alive, nothing deferred to the later phases
        uint32_t m0, m1;
        asm("set.ne.u32.f32 %0, %1, %2;" : "=r"(m0) : "f"(r0), "f"(a.nodata));
        asm("set.ne.u32.f32 %0, %1, %2;" : "=r"(m1) : "f"(r1), "f"(a.nodata));
        const uint32_t b0 = __float_as_uint(r0), b1 = __float_as_uint(r1), pb = __float_as_uint(piv);
        t.xs[i] = __uint_as_float((b0 & m0) | (0x7fc00000u & ~m0));
        t.xs[i + 1] = __uint_as_float((b1 & m1) | (0x7fc00000u & ~m1));
        const float x0 = __uint_as_float((b0 & m0) | (pb & ~m0));
        const float x1 = __uint_as_float((b1 & m1) | (pb & ~m1));
        asm("min.f32 %0, %0, %1, %2;" : "+f"(mn) : "f"(x0), "f"(x1));
        asm("max.f32 %0, %0, %1, %2;" : "+f"(mx) : "f"(x0), "f"(x1));
        unsigned long long xx, d, d2;
        asm("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(x0), "f"(x1));
        asm("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(xx), "l"(pp));
        asm("mul.f32x2 %0, %1, %1;" : "=l"(d2) : "l"(d));
        asm("add.f32x2 %0, %0, %1;" : "+l"(s1) : "l"(d));
        asm("add.f32x2 %0, %0, %1;" : "+l"(s2) : "l"(d2));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(s3) : "l"(d2), "l"(d));
    }
    const float S1 = __uint_as_float(static_cast<uint32_t>(s1)) + __uint_as_float(static_cast<uint32_t>(s1 >> 32));
    const float S2 = __uint_as_float(static_cast<uint32_t>(s2)) + __uint_as_float(static_cast<uint32_t>(s2 >> 32));
    const float S3 = __uint_as_float(static_cast<uint32_t>(s3)) + __uint_as_float(static_cast<uint32_t>(s3 >> 32));
#else
    float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f}, s3[2] = {0.f, 0.f};
    for (int i = 0; i < IPT; ++i) {
        const float r = src[i * NW * COLS];
        const bool v = valid(r, a.nodata);
        const float x = v ? r : piv;
        t.xs[i] = v ? r : qnan();
        mn = fminf(mn, x);
        mx = fmaxf(mx, x);
        const float d = x - piv;
        const float d2 = d * d;
        s1[i & 1] += d;
        s2[i & 1] += d2;
        s3[i & 1] = fmaf(d2, d, s3[i & 1]);
    }
    const float S1 = s1[0] + s1[1], S2 = s2[0] + s2[1], S3 = s3[0] + s3[1];
#endif
    uint32_t* q = s.part + w * COLS + c;
    q[0 * NW * COLS] = f2u(S1);
    q[1 * NW * COLS] = f2u(S2);
    q[2 * NW * COLS] = f2u(S3);
    // an empty column keeps piv = 0 in mn / mx: harmless, n == 0 decides its planes
    atomic_min_u32(s.mnmx + c, okey(mn));
    rsel::atomic_max_u32(s.mnmx + COLS + c, okey(mx));
}

// ------------------------------------------------------------------ phase C1
// range and bin map of the thread's column (every thread, two loads); warps
// 0..2 add the NW partials of one power sum in float64, in the register-tile
// kernel's order (groups of 8, then the groups), so that the moment planes of
// the two kernels agree bit for bit
template <int NW, int IPT>
RS_HD void phase_c1(Thr<IPT>& t, const Smem& s, int w, int c) {
    static_assert(NW % 8 == 0, "NW must be a multiple of 8");
    const float mn = okey_inv(s.mnmx[c]), mx = okey_inv(s.mnmx[COLS + c]);
    t.lo = mn;
    t.hi = mx;
    const float d = mx - mn;
    float scale = static_cast<float>(NBINS - 1) / d;
    t.dead = 0;
    if (!(scale < pinf()) || !(scale > 0.f)) {
        // constant column (d == 0): nothing to bin.  Otherwise the range over-
        // or underflows float32: the tile goes to the fallback kernel
        if (mx > mn) t.dead = 1;
        scale = 0.f;
    }
    t.scale = scale;
    if (w < 3) {
        const uint32_t* p = s.part + (w * NW) * COLS + c;
        double tot = 0.0;
        for (int sub = 0; sub < NW / 8; ++sub) {
            double acc = 0.0;
            for (int j = 0; j < 8; ++j) acc += static_cast<double>(u2f(p[(sub * 8 + j) * COLS]));
            tot += acc;
        }
        s.msum[w * COLS + c] = tot;
    }
}

// ------------------------------------------------------------------ phase B
template <int NW, int IPT>
RS_HD void phase_b(const Thr<IPT>& t, const Smem& s, int c) {
    uint32_t* cells = s.cells + c;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const int b = bin_of(t.xs[i], t.lo, t.scale);
#ifdef __CUDA_ARCH__
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(cells + b * COLS))) : "memory");
#else
        __atomic_fetch_add(cells + b * COLS, 1u, __ATOMIC_RELAXED);
#endif
    }
}

// ------------------------------------------------------------------ phase S1
template <int NW>
RS_HD void phase_s1(const Smem& s, int w, int c) {
    constexpr int CH = (NBINS + NW - 1) / NW;
    const int b0 = w * CH;
    uint32_t tot = 0;
#pragma unroll 8
    for (int i = 0; i < CH; ++i)
        if (b0 + i < NBINS) tot += s.cells[(b0 + i) * COLS + c];
    s.chunk[w * COLS + c] = tot;
    // `part` (read in C1, before the last barrier) is dead: its words become the lists
    if (w < NLMAX) s.lists[w * LSTRIDE + c] = 0u;
}

// ------------------------------------------------------------------ phase L
// np.quantile(method='linear'): h = (n - 1) q, ranks floor(h) and floor(h) + 1
// (clamped so that both exist).  Warp j looks rank (j & 1) of quantile j >> 1 up
// for its 32 columns: the chunk whose cumulative total passes the rank, then the
// bin inside that chunk, from the raw counts (marks of other warps are masked
// off).  The bin is claimed for list j by compare-and-swap; when another rank of
// the column already claimed it, that list is shared.
RS_HD int rank_floor(int n, double q) {
    const double h = dmul(static_cast<double>(n) - 1.0, q);
    int l0 = static_cast<int>(floor(h));
    const int top = n - 2;
    l0 = l0 > top ? top : l0;
    l0 = l0 < 0 ? 0 : l0;
    return l0;
}

template <int NW, int IPT>
RS_HD void phase_locate(const Thr<IPT>& t, const Smem& s, const Args& a, int w, int c) {
    constexpr int CH = (NBINS + NW - 1) / NW;
    uint32_t tot[NW];
    uint32_t n = 0;
#pragma unroll
    for (int g = 0; g < NW; ++g) {
        tot[g] = s.chunk[g * COLS + c];
        n += tot[g];
    }
    if (w == 0) s.ncol[c] = n;
    for (int j = w; j < 2 * a.nq && j < NLMAX; j += NW) {        // warp-uniform
        uint32_t info = 0u;
        if (n >= 2u && t.hi > t.lo && !t.dead) {
            const int pl = a.qplane[j >> 1];
            const uint32_t r = static_cast<uint32_t>(rank_floor(static_cast<int>(n), a.q[pl]) + (j & 1));
            uint32_t cum = 0, base = 0;
            int g = 0;
#pragma unroll
            for (int gg = 0; gg < NW; ++gg) {
                cum += tot[gg];
                if (cum <= r) { g = gg + 1; base = cum; }
            }
            g = g < NW ? g : NW - 1;                     // r < n: never taken
            const int b0 = g * CH;
            uint32_t cumb = base, start = base;
            int bsel = b0;
#pragma unroll 8
            for (int i = 0; i < CH; ++i) {
                const int b = b0 + i;
                const uint32_t cnt = b < NBINS ? (ld_racy(s.cells + b * COLS + c) & COUNT_MASK) : 0u;
                cumb += cnt;
                if (cumb <= r && b + 1 < NBINS) { bsel = b + 1; start = cumb; }
            }
            uint32_t* cell = s.cells + bsel * COLS + c;
            const uint32_t cnt = ld_racy(cell) & COUNT_MASK;
            const uint32_t old = atomic_cas_u32(cell, cnt, cnt | (static_cast<uint32_t>(j + 1) << MARK_SHIFT));
            const uint32_t lid = (old >> MARK_SHIFT) ? (old >> MARK_SHIFT) - 1u : static_cast<uint32_t>(j);
            const uint32_t kind = bsel == 0 ? 1u : bsel == NBINS - 1 ? 2u : 0u;
            info = RI_VALID | (kind << 26) | (lid << 22) | (cnt << 11) | (r - start);
        }
        s.rinfo[j * COLS + c] = info;
    }
}

// ------------------------------------------------------------------ phase X
// the values of claimed bins -> their lists.  Values equal to the column minimum
// or maximum (the clamp atoms, possibly hundreds) are not copied: the first bin
// holds the minimum atoms, then its regular values; the last bin its regular
// values, then the maximum atoms -- their number is the bin's population minus
// the list's length.
template <int NW, int IPT>
RS_HD void phase_x(const Thr<IPT>& t, const Smem& s, int c) {
    const uint32_t* cells = s.cells + c;
    uint32_t* lists = s.lists + c;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const float x = t.xs[i];
        const int b = bin_of(x, t.lo, t.scale);
        const uint32_t mk = cells[b * COLS] >> MARK_SHIFT;
        if (mk != 0u) {
            if (x > t.lo && x < t.hi) {
                uint32_t* l = lists + (mk - 1u) * LSTRIDE;
                const uint32_t slot = rsel::atomic_add(l, 1u);
                if (slot < static_cast<uint32_t>(CAP)) l[(slot + 1u) * COLS] = f2u(x);
            }
        }
    }
}

// ------------------------------------------------------------------ phase O
// list -> registers (padded with +INF) -> sorting network; `K` is the smallest
// of 8 / 16 / 32 that holds the longest list among the warp's columns
template <int K>
RS_HD void load_sorted(const uint32_t* l, int m, float (&v)[K]) {
#pragma unroll
    for (int i = 0; i < K; ++i) v[i] = i < m ? u2f(l[(i + 1) * COLS]) : pinf();
    rsel::sort_net<K>(v);
}

template <int K>
RS_HD float list_min(const uint32_t* l, int m) {
    float r = pinf();
#pragma unroll
    for (int i = 0; i < K; ++i) r = fminf(r, i < m ? u2f(l[(i + 1) * COLS]) : pinf());
    return r;
}

struct RankRec {
    int o, cnt, lid, kind;
};
RS_HD RankRec decode_rank(uint32_t info) {
    RankRec r;
    r.o = static_cast<int>(info & 0x7ffu);
    r.cnt = static_cast<int>((info >> 11) & 0x7ffu);
    r.lid = static_cast<int>((info >> 22) & 0xfu);
    r.kind = static_cast<int>((info >> 26) & 0x3u);
    return r;
}

// value at offset `idx` of a bin whose regular values are v[0..m) in order:
// kind 1 (first bin) has cnt - m minimum atoms in front, the last bin its
// maximum atoms behind (an inner bin has none: idx < m always)
template <int K>
RS_HD float bin_value(const float (&v)[K], int m, int idx, float lo, float hi) {
    int k = idx < 0 ? 0 : idx;
    k = k > K - 1 ? K - 1 : k;
    const float x = rsel::pick<K>(v, k);
    return idx < 0 ? lo : (idx < m ? x : hi);
}

template <int K>
RS_HD void rank_values(const Smem& s, const RankRec& r0, const RankRec& r1, int m0, int m1, float lo,
                       float hi, int c, float* f0, float* f1) {
    float v[K];
    load_sorted<K>(s.lists + r0.lid * LSTRIDE + c, m0, v);
    const int shift = r0.kind == 1 ? r0.cnt - m0 : 0;
    *f0 = bin_value<K>(v, m0, r0.o - shift, lo, hi);
    const float same = bin_value<K>(v, m0, r0.o + 1 - shift, lo, hi);
    // another bin: rank l0 + 1 is its first value (the smallest regular value,
    // or -- last bin without regular values -- the maximum)
    const float mn1 = list_min<K>(s.lists + r1.lid * LSTRIDE + c, m1);
    const float other = m1 > 0 ? mn1 : hi;
    *f1 = (r1.lid == r0.lid) ? same : other;
}

// quantile plane `pl` (the k-th quantile plane) of column c; *flag is raised
// when a list overflowed
template <int NW, int IPT>
RS_HD float quantile_plane(const Thr<IPT>& t, const Smem& s, const Args& a, int k, int pl, int c,
                           bool* flag) {
    const int n = static_cast<int>(s.ncol[c]);
    const uint32_t i0 = s.rinfo[(2 * k) * COLS + c], i1 = s.rinfo[(2 * k + 1) * COLS + c];
    const bool have = (i0 & RI_VALID) != 0u && (i1 & RI_VALID) != 0u;
    RankRec r0 = decode_rank(i0), r1 = decode_rank(i1);
    if (!have) { r0.lid = r1.lid = 0; r0.o = r1.o = 0; r0.cnt = r1.cnt = 0; r0.kind = r1.kind = 0; }
    int m0 = have ? static_cast<int>(s.lists[r0.lid * LSTRIDE + c]) : 0;
    int m1 = have ? static_cast<int>(s.lists[r1.lid * LSTRIDE + c]) : 0;
    if (m0 > CAP || m1 > CAP) { *flag = true; m0 = m0 > CAP ? CAP : m0; m1 = m1 > CAP ? CAP : m1; }
    int need = m0 > m1 ? m0 : m1;
#ifdef __CUDA_ARCH__
    need = __reduce_max_sync(0xffffffffu, need);
#endif
    float f0, f1;
    if (need <= 8) rank_values<8>(s, r0, r1, m0, m1, t.lo, t.hi, c, &f0, &f1);
    else if (need <= 16) rank_values<16>(s, r0, r1, m0, m1, t.lo, t.hi, c, &f0, &f1);
    else rank_values<32>(s, r0, r1, m0, m1, t.lo, t.hi, c, &f0, &f1);
    if (n == 0) return qnan();
    if (!(t.hi > t.lo)) return t.lo;
    if (!have) return qnan();            // dead column (flagged by the caller) -- n >= 2 otherwise
    // NumPy _lerp, operation by operation in float64 (rsel::quantile_value)
    const double q = a.q[pl];
    const double dn = static_cast<double>(n);
    const double h = dmul(dn - 1.0, q);
    const double fl = floor(h);
    const double tt = h - fl;
    const double x0 = static_cast<double>(f0), x1 = static_cast<double>(f1);
    const double d = x1 - x0;
    const double rq = (tt >= 0.5) ? dsub(x1, dmul(d, 1.0 - tt)) : dadd(x0, dmul(d, tt));
    float r = static_cast<float>(rq);
    if (h >= dn - 1.0) r = t.hi;
    if (h < 0.0) r = t.lo;
    return r;
}

// moment planes: float64 finals from the float64 sums of the float32 partials
// (rsel::plane_value, same operations in the same order)
template <int NW, int IPT>
RS_HD float moment_plane(const Thr<IPT>& t, const Smem& s, int kind, int c) {
    const int n = static_cast<int>(s.ncol[c]);
    if (n == 0) return qnan();
    const double S1 = s.msum[0 * COLS + c], S2 = s.msum[1 * COLS + c], S3 = s.msum[2 * COLS + c];
    const double dn = static_cast<double>(n);
    const double delta = S1 / dn;
    const double mean = static_cast<double>(t.piv) + delta;
    const bool allsame = !(t.hi > t.lo);
    double m2 = S2 - S1 * delta;
    double m3 = S3 - 3.0 * delta * S2 + 2.0 * S1 * delta * delta;
    if (allsame || S2 == 0.0) { m2 = 0.0; m3 = 0.0; }
    if (m2 < 0.0) m2 = 0.0;
    const double nanv = static_cast<double>(qnan());
    double r;
    if (kind == PK_MEAN) {
        r = allsame ? static_cast<double>(t.lo) : mean;
    } else if (kind == PK_VAR) {
        r = n > 1 ? m2 / (dn - 1.0) : nanv;
    } else if (kind == PK_STD) {
        r = n > 1 ? sqrt(m2 / (dn - 1.0)) : nanv;
    } else if (kind == PK_COEFVAR) {
        const double mu = allsame ? static_cast<double>(t.lo) : mean;
        r = n > 1 ? sqrt(m2 / (dn - 1.0)) / mu * 100.0 : nanv;
    } else {   // PK_SKEW: adjusted Fisher-Pearson G1 (pd.Series.skew)
        if (n < 3) r = nanv;
        else if (m2 == 0.0) r = 0.0;
        else r = dn * sqrt(dn - 1.0) / (dn - 2.0) * (m3 / (m2 * sqrt(m2)));
    }
    return static_cast<float>(r);
}

// plane pl of column c (warp-uniform pl): kq = index of pl among the quantile planes
template <int NW, int IPT>
RS_HD float plane_value(const Thr<IPT>& t, const Smem& s, const Args& a, int pl, int c, bool* flag) {
    const int kind = a.kind[pl];
    if (kind == PK_QUANT) {
        int k = 0;
        for (int i = 0; i < a.nq; ++i)
            if (a.qplane[i] == pl) k = i;
        return quantile_plane<NW, IPT>(t, s, a, k, pl, c, flag);
    }
    return moment_plane<NW, IPT>(t, s, kind, c);
}

}  // namespace ssel
