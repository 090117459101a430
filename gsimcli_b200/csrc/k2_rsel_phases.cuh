// k2_rsel_phases.cuh -- the per-thread phases of the K2 "register tile" kernel
// (k2_grid_rsel.cu): GridFiles.stats cell loop, GSIMCLI/tools/grid.py:662-701,
// for 32 < R <= 1024.
//
// One CTA owns a tile of 32 adjacent grid nodes ("columns") x R realizations.
// LANE = COLUMN: thread (warp w, lane c) holds rows w, w + NW, w + 2 NW ... of
// column c in registers (IPT values), loaded with coalesced 128-byte row reads.
// Every shared-memory structure is laid out [index][32 columns], so a warp
// instruction touches bank c from lane c: histogram updates, the scatter and all
// look-ups are bank-conflict free by construction, and a column full of equal
// values (DSS clamps to [datamin, datamax]) costs nothing extra.
//
// Phases, separated by CTA barriers (each is a plain per-thread function, which
// is what lets tests/emul run the very same code on the host, thread by thread):
//   pivot   first valid value of the column (shift of the one-pass moments)
//   A       nodata/NaN -> NaN; count, min, max, shifted power sums (float32,
//           two interleaved accumulators = the packed f32x2 pipeline)
//   combine partial sums of the NW warps -> float64; [lo, hi] -> linear bin map
//   B       histogram: one conflict-free RED per value into cells[bin][c]
//   scan    cells -> exclusive prefix (= first slot of every bin)
//   C       scatter: slot = ATOMS(cells[bin][c], 1); sortbuf[slot][c] = x.  The
//           column is now ordered by bin; values equal to the column minimum /
//           maximum (the clamp atoms) are not moved, only counted.
//   finish  rank look-up = binary search over the bin ends + selection inside
//           ONE bin (a handful of values); mode = pair comparison inside bins
//           that hold at least as many values as the best multiplicity so far
//   out     planes -> global memory; cells back to zero
// Exactness: bins are value-ordered and every comparison is on the original
// float32, so ranks, quantile operands and the mode are exact.  A column whose
// bin map is useless (range overflow, one bin with more than RS_BIGBIN values
// that is needed) raises the tile's flag and the tile is redone by the
// counting-sort kernel (k2_grid_csort.cu) -- never a guess.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define RS_HD __device__ __forceinline__
#else
#define RS_HD static inline
#endif

#ifndef GSB_MAX_PLANES
#define GSB_MAX_PLANES 64
#endif

namespace rsel {

constexpr int COLS = 32;
constexpr int NBINS = 512;
constexpr int TRASH = NBINS;            // cell of the missing values
constexpr int NCELLS = NBINS + 1;
constexpr int RS_BIGBIN = 48;           // a needed bin larger than this -> fallback
constexpr uint32_t MAGIC_BITS = 0x4B400000u;   // 1.5 * 2^23

// plane kinds (mirror gsb_common.cuh)
enum { PK_MEAN = 0, PK_QUANT = 1, PK_SKEW = 2, PK_VAR = 3, PK_STD = 4, PK_COEFVAR = 5, PK_MODE = 6 };

// per-column words of `colinfo` ([CI_WORDS][32])
enum { CI_N0TOT = 0, CI_START511 = 1, CI_N511TOT = 2, CI_HINT = 3, CI_FLAG = 4, CI_WORDS = 5 };

struct Args {
    int R;
    int nplanes;
    int nq;                 // quantile planes
    float nodata;
    unsigned char kind[GSB_MAX_PLANES];
    unsigned char qplane[GSB_MAX_PLANES];   // plane index of the k-th quantile plane
    double q[GSB_MAX_PLANES];
};

// pointers into the CTA's shared memory
struct Smem {
    float* sortbuf;               // [rows][32]; aliased by part / chunk before phase C
    uint32_t* part;               // [6][NW][32]   (alias of sortbuf)
    uint32_t* chunk;              // [NW][32]      (alias of sortbuf)
    uint32_t* cells;              // [NCELLS][32]
    double* part2;                // [6][S][32]
    float* res;                   // [GSB_MAX_PLANES][32]
    float* pivc;                  // [NW][32]
    uint32_t* colinfo;            // [CI_WORDS][32]
    unsigned long long* modekey;  // [32]
};

template <int IPT>
struct Thr {
    float xs[IPT];
    float piv, lo, hi, scale;
    int n;          // valid values of the column
    int dead;       // column not handled here (range overflow): the tile is flagged
};

// ------------------------------------------------------------------ primitives
RS_HD uint32_t f2u(float x) { uint32_t u; memcpy(&u, &x, 4); return u; }
RS_HD float u2f(uint32_t u) { float x; memcpy(&x, &u, 4); return x; }
RS_HD float qnan() { return u2f(0x7fc00000u); }
RS_HD float pinf() { return u2f(0x7f800000u); }

// bn.replace(arr, nodata, nan) + NaN dropping (grid.py:681): false for NaN
RS_HD bool valid(float x, float nodata) { return (x == x) && (x != nodata); }

// (host forms: real atomics, so that the emulations may also run the phases with
// one OS thread per CUDA thread -- tests/emul/k2_ssel_threads.cpp)
RS_HD uint32_t atomic_add(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
    return atomicAdd(p, v);
#else
    return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
#endif
}
RS_HD void atomic_max_u32(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
    atomicMax(p, v);
#else
    uint32_t o = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
#endif
}
RS_HD void atomic_max_u64(unsigned long long* p, unsigned long long v) {
#ifdef __CUDA_ARCH__
    atomicMax(p, v);
#else
    unsigned long long o = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
#endif
}
RS_HD double dmul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
RS_HD double dadd(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
RS_HD double dsub(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}

// value -> bin: round((x - lo) * scale) clamped to [0, NBINS-1]; NaN -> TRASH.
// The float sum lands on MAGIC + integer, so the integer sits in the low
// mantissa bits; one fused add-min-relu does the subtraction and both clamps.
RS_HD int bin_of(float x, float lo, float scale) {
    const float t = fmaf(x - lo, scale, u2f(MAGIC_BITS));
#ifdef __CUDA_ARCH__
    return __viaddmin_s32_relu(static_cast<int>(f2u(t)), -static_cast<int>(MAGIC_BITS), TRASH);
#else
    long long b = static_cast<long long>(static_cast<int32_t>(f2u(t))) - static_cast<long long>(MAGIC_BITS);
    if (b > TRASH) b = TRASH;
    if (b < 0) b = 0;
    return static_cast<int>(b);
#endif
}

// order-preserving key of a float (for the 64-bit mode record)
RS_HD uint32_t okey(float v) {
    const uint32_t b = f2u(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
RS_HD float okey_inv(uint32_t k) {
    return u2f((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
RS_HD unsigned long long mode_key(uint32_t mult, float v) {
    return (static_cast<unsigned long long>(mult) << 32) | (0xffffffffu - okey(v));
}

// ------------------------------------------------------------------ networks
// ---- in-bin selection without data-dependent loops (lane = column: a loop whose
// trip count is the bin's population would make the warp pay the largest bin of
// its 32 columns squared).  The bin's values, padded with +INF, go through a
// fixed odd-even merge sorting network in registers; the wanted order statistic
// is picked with a select tree.
RS_HD void cswap(float& a, float& b) {
    const float lo = fminf(a, b), hi = fmaxf(a, b);
    a = lo;
    b = hi;
}

template <int K>
RS_HD void sort_net(float (&v)[K]) {
    // Batcher's odd-even merge sort, K a power of two (19 exchanges for 8, 63 for 16)
#pragma unroll
    for (int p = 1; p < K; p <<= 1)
#pragma unroll
        for (int k = p; k >= 1; k >>= 1)
#pragma unroll
            for (int j = k % p; j + k < K; j += 2 * k)
#pragma unroll
                for (int i = 0; i < k; ++i)
                    if (i + j + k < K && (i + j) / (2 * p) == (i + j + k) / (2 * p))
                        cswap(v[i + j], v[i + j + k]);
}

template <int K>
RS_HD float pick(const float (&v)[K], int o) {
    // v[o] for 0 <= o < K by a binary select tree on the bits of o
    float t[K];
#pragma unroll
    for (int i = 0; i < K; ++i) t[i] = v[i];
#pragma unroll
    for (int step = 1; step < K; step <<= 1) {
        const bool bit = (o & step) != 0;
#pragma unroll
        for (int i = 0; i + step < K; i += 2 * step) t[i] = bit ? t[i + step] : t[i];
    }
    return t[0];
}

// ------------------------------------------------------------------ pivot
// first valid value among the thread's own rows -> pivc[w][c]
template <int NW, int IPT>
RS_HD void phase_pivot_write(const Thr<IPT>& t, const Smem& s, const Args& a, int w, int c) {
    float cand = qnan();
    if (valid(t.xs[0], a.nodata)) {
        cand = t.xs[0];
    } else {
#pragma unroll
        for (int i = IPT - 1; i >= 1; --i)
            if (valid(t.xs[i], a.nodata)) cand = t.xs[i];
    }
    s.pivc[w * COLS + c] = cand;
}

// Pivot of the column = the lower MEDIAN of the candidates (one element per warp:
// the first valid value of warps 0..7).  It must be an element (it stands in for
// missing values in min / max) and close to the mean: the float32 shifted sums
// lose |pivot - mean| / sigma digits in the third moment.  (First-valid-element
// pivots gave 5e-5 skewness errors on gamma columns; the candidate nearest to the
// candidates' MEAN, used until the CPU fuzz of round 2, was dragged 4-5 sigma out
// by one or two tail values when only a few candidates exist -- every second
// realization missing left four -- and cost 1e-4 in the skewness.)
// Every thread of the column computes the same value from the same words.
// `get(w)` = candidate of warp w (NaN: none).  The first 8 valid candidates in warp
// order take part: warps 0..7 as a rule, later warps only when some of those have
// no valid value at all.
template <int NW, class F>
RS_HD float pivot_of_candidates(F get) {
    float v[8];
    int cnt = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const float x = get(w);
        const bool ok = x == x;
        v[w] = ok ? x : pinf();
        cnt += ok ? 1 : 0;
    }
    if (cnt < 8) {                                  // rare: top up from the other warps
        for (int w = 8; w < NW && cnt < 8; ++w) {
            const float x = get(w);
            if (x == x) {
                bool placed = false;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const bool take = !placed && !(v[i] < pinf());
                    v[i] = take ? x : v[i];
                    placed = placed || take;
                }
                ++cnt;
            }
        }
    }
    if (cnt == 0) return 0.f;
    sort_net<8>(v);
    return pick<8>(v, (cnt - 1) >> 1);
}

template <int NW, int IPT>
RS_HD void phase_pivot_read(Thr<IPT>& t, const Smem& s, int c) {
    const float* pc = s.pivc + c;
    t.piv = pivot_of_candidates<NW>([pc](int w) { return pc[w * COLS]; });
}

// ------------------------------------------------------------------ phase A
// the pivot is an element of the column, so it stands in for missing values in
// the sums (d == 0), the minimum and the maximum
template <int NW, int IPT>
RS_HD void phase_a(Thr<IPT>& t, const Smem& s, const Args& a, int w, int c) {
    static_assert(IPT % 2 == 0, "IPT must be even");
    const float piv = t.piv;
    float mn = piv, mx = piv;
    int n = 0;
#ifdef __CUDA_ARCH__
    unsigned long long s1 = 0, s2 = 0, s3 = 0, pp;
    asm("mov.b64 %0, {%1, %1};" : "=l"(pp) : "f"(piv));
#pragma unroll
    for (int i = 0; i < IPT; i += 2) {
        const bool v0 = valid(t.xs[i], a.nodata), v1 = valid(t.xs[i + 1], a.nodata);
        const float x0 = v0 ? t.xs[i] : piv, x1 = v1 ? t.xs[i + 1] : piv;
        t.xs[i] = v0 ? t.xs[i] : qnan();
        t.xs[i + 1] = v1 ? t.xs[i + 1] : qnan();
        n += (v0 ? 1 : 0) + (v1 ? 1 : 0);
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
        const bool v = valid(t.xs[i], a.nodata);
        const float x = v ? t.xs[i] : piv;
        t.xs[i] = v ? t.xs[i] : qnan();
        n += v ? 1 : 0;
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
    uint32_t* p = s.part + w * COLS + c;
    p[0 * NW * COLS] = static_cast<uint32_t>(n);
    p[1 * NW * COLS] = f2u(S1);
    p[2 * NW * COLS] = f2u(S2);
    p[3 * NW * COLS] = f2u(S3);
    p[4 * NW * COLS] = f2u(mn);
    p[5 * NW * COLS] = f2u(mx);
}

// ------------------------------------------------------------------ combine
// stage 1: warp (k, sub) adds 8 partials of quantity k in float64
template <int NW>
RS_HD void phase_combine1(const Smem& s, int w, int c) {
    constexpr int S = NW / 8;
    if (w >= 6 * S) return;
    const int k = w / S, sub = w % S;
    const uint32_t* p = s.part + (k * NW + sub * 8) * COLS + c;
    double acc;
    if (k == 0) {
        acc = 0.0;
        for (int j = 0; j < 8; ++j) acc += static_cast<double>(p[j * COLS]);
    } else if (k <= 3) {
        acc = 0.0;
        for (int j = 0; j < 8; ++j) acc += static_cast<double>(u2f(p[j * COLS]));
    } else if (k == 4) {
        float m = u2f(p[0]);
        for (int j = 1; j < 8; ++j) m = fminf(m, u2f(p[j * COLS]));
        acc = static_cast<double>(m);
    } else {
        float m = u2f(p[0]);
        for (int j = 1; j < 8; ++j) m = fmaxf(m, u2f(p[j * COLS]));
        acc = static_cast<double>(m);
    }
    s.part2[(k * S + sub) * COLS + c] = acc;
}

// stage 2 (every thread, redundantly -- no barrier, no serial warp): count,
// range and bin map of the thread's column
template <int NW, int IPT>
RS_HD void phase_combine2(Thr<IPT>& t, const Smem& s, int w, int c) {
    constexpr int S = NW / 8;
    double n = 0.0;
    float mn = pinf(), mx = -pinf();
    for (int sub = 0; sub < S; ++sub) {
        n += s.part2[(0 * S + sub) * COLS + c];
        mn = fminf(mn, static_cast<float>(s.part2[(4 * S + sub) * COLS + c]));
        mx = fmaxf(mx, static_cast<float>(s.part2[(5 * S + sub) * COLS + c]));
    }
    t.n = static_cast<int>(n);
    t.lo = mn;
    t.hi = mx;
    const float d = mx - mn;
    float scale = static_cast<float>(NBINS - 1) / d;
    t.dead = 0;
    if (!(scale < pinf()) || !(scale > 0.f)) {
        // constant / empty column (d == 0): nothing to bin.  Otherwise the range
        // over- or underflows float32: the tile goes to the fallback kernel
        if (t.n > 0 && mx > mn) t.dead = 1;
        scale = 0.f;
    }
    t.scale = scale;
    if (w == 0) {
        s.colinfo[CI_FLAG * COLS + c] = t.dead ? 1u : 0u;
        s.colinfo[CI_HINT * COLS + c] = 0u;
        s.modekey[c] = 0ull;
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

// ------------------------------------------------------------------ scan
template <int NW>
RS_HD void phase_scan1(const Smem& s, int w, int c) {
    constexpr int CH = (NBINS + NW - 1) / NW;
    const int b0 = w * CH;
    uint32_t tot = 0;
#pragma unroll 8
    for (int i = 0; i < CH; ++i)
        if (b0 + i < NBINS) tot += s.cells[(b0 + i) * COLS + c];
    s.chunk[w * COLS + c] = tot;
}

template <int NW>
RS_HD void phase_scan2(const Smem& s, int w, int c) {
    constexpr int CH = (NBINS + NW - 1) / NW;
    const int b0 = w * CH;
    uint32_t run = 0;
#pragma unroll
    for (int g = 0; g < NW; ++g) {            // all loads in flight at once; w is warp-uniform
        const uint32_t v = s.chunk[g * COLS + c];
        run += g < w ? v : 0u;
    }
#pragma unroll 8
    for (int i = 0; i < CH; ++i) {
        const int b = b0 + i;
        if (b < NBINS) {
            const uint32_t cnt = s.cells[b * COLS + c];
            s.cells[b * COLS + c] = run;
            if (b == 0) s.colinfo[CI_N0TOT * COLS + c] = cnt;
            if (b == NBINS - 1) {
                s.colinfo[CI_START511 * COLS + c] = run;
                s.colinfo[CI_N511TOT * COLS + c] = cnt;
            }
            run += cnt;
        }
    }
}

// ------------------------------------------------------------------ phase C
// only the values strictly between the column minimum and maximum move; the
// atoms at both ends are recovered from the bin totals (n0tot, n511tot)
template <int NW, int IPT>
RS_HD void phase_c(const Thr<IPT>& t, const Smem& s, int c) {
    uint32_t* cells = s.cells + c;
    float* buf = s.sortbuf + c;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const float x = t.xs[i];
        if (x > t.lo && x < t.hi) {
            const int b = bin_of(x, t.lo, t.scale);
            const uint32_t slot = atomic_add(cells + b * COLS, 1u);
            buf[slot * COLS] = x;
        }
    }
}

// ------------------------------------------------------------------ finish
// Sorted order of the column's n valid values:
//   [n_lo x lo] [bin 0 regular] [bin 1] ... [bin 510] [bin 511 regular] [n_hi x hi]
// with bin b's regular values in slots [start_b, end_b): start_0 = 0,
// start_1 = n0tot, start_b = end_{b-1}; end_b = cells[b] after phase C.
struct ColView {
    const uint32_t* cells;   // + c
    const float* buf;        // + c
    int n, n_lo, n_hi, reg0, n0tot, start511, end511;
    float lo, hi;
};

template <int IPT>
RS_HD ColView make_view(const Thr<IPT>& t, const Smem& s, int c) {
    ColView v;
    v.cells = s.cells + c;
    v.buf = s.sortbuf + c;
    v.n = t.n;
    v.lo = t.lo;
    v.hi = t.hi;
    v.n0tot = static_cast<int>(s.colinfo[CI_N0TOT * COLS + c]);
    v.reg0 = static_cast<int>(v.cells[0]);
    v.n_lo = v.n0tot - v.reg0;
    v.start511 = static_cast<int>(s.colinfo[CI_START511 * COLS + c]);
    v.end511 = static_cast<int>(v.cells[(NBINS - 1) * COLS]);
    v.n_hi = static_cast<int>(s.colinfo[CI_N511TOT * COLS + c]) - (v.end511 - v.start511);
    if (!(t.hi > t.lo)) {          // constant column: every value is the "lo" atom
        v.n_lo = t.n;
        v.n_hi = 0;
        v.reg0 = 0;
    }
    return v;
}

// o-th (and, when o + 1 < cnt, (o+1)-th) smallest of the cnt <= K values of a bin
template <int K>
RS_HD void select_net(const float* p, int cnt, int o, float* f0, float* f1) {
    float v[K];
#pragma unroll
    for (int i = 0; i < K; ++i) v[i] = i < cnt ? p[i * COLS] : pinf();
    sort_net<K>(v);
    *f0 = pick<K>(v, o);
    const int o1 = o + 1 < K ? o + 1 : K - 1;
    *f1 = pick<K>(v, o1);
}

// smallest of the cnt <= RS_BIGBIN values of a bin (first value of the next rank's bin)
RS_HD float bin_min(const float* p, int cnt) {
    float m = pinf();
    for (int i = 0; i < cnt; ++i) m = fminf(m, p[i * COLS]);
    return m;
}

// generic in-bin selection (cnt > 16: rare -- a dense spot of the distribution)
RS_HD void select_loop(const float* p, int cnt, int o, float* f0, float* f1) {
    for (int i = 0; i < cnt; ++i) {
        const float x = p[i * COLS];
        int rank = 0;
        for (int j = 0; j < cnt; ++j) {
            const float y = p[j * COLS];
            rank += (y < x || (y == x && j < i)) ? 1 : 0;
        }
        if (rank == o) *f0 = x;
        if (rank == o + 1) *f1 = x;
    }
}

// where sorted rank r lives: an atom (returns true, *val set) or slot o of the
// bin segment [start, start + cnt)
RS_HD bool locate(const ColView& v, int r, float* val, int* start, int* cnt, int* o) {
    if (r < v.n_lo) { *val = v.lo; return true; }
    if (r < v.n_lo + v.reg0) { *start = 0; *cnt = v.reg0; *o = r - v.n_lo; return false; }
    if (r >= v.end511) { *val = v.hi; return true; }
    // smallest b in [1, NBINS-1] with end_b > r: 9 uniform halving steps
    int lo = 1;
#pragma unroll
    for (int step = NBINS / 2; step >= 1; step >>= 1) {
        const int mid = lo + step - 1;          // candidate: is end_mid <= r ?
        if (mid < NBINS - 1 && static_cast<int>(v.cells[mid * COLS]) <= r) lo = mid + 1;
    }
    const int s0 = (lo == 1) ? v.n0tot : static_cast<int>(v.cells[(lo - 1) * COLS]);
    *start = s0;
    *cnt = static_cast<int>(v.cells[lo * COLS]) - s0;
    *o = r - s0;
    return false;
}

// values at sorted ranks r and r + 1 (g1 only when want1; r + 1 < n required).
// `wide` (warp-uniform on the device) says some column of the warp needs the
// 16-value network.  Returns false when a bin too large to select from is
// involved (the caller flags the tile).
RS_HD bool rank_pair(const ColView& v, int r, bool want1, bool wide, float* g0, float* g1) {
    int start = 0, cnt = 0, o = 0;
    const bool atom0 = locate(v, r, g0, &start, &cnt, &o);
    bool same_bin = false;
    if (!atom0) {
        const float* p = v.buf + start * COLS;
        float h1 = 0.f;
        if (cnt > RS_BIGBIN) {
            // too many values to select from -- unless they are all equal (heavy
            // ties: quantised data put hundreds of equal values into one bin),
            // in which case every rank of the bin is that value
            float mn = p[0], mx = p[0];
            for (int i = 1; i < cnt; ++i) {
                const float x = p[i * COLS];
                mn = fminf(mn, x);
                mx = fmaxf(mx, x);
            }
            if (mn != mx) return false;
            *g0 = mn;
            h1 = mn;
        } else if (cnt > 16) select_loop(p, cnt, o, g0, &h1);
        else if (wide || cnt > 8) select_net<16>(p, cnt, o, g0, &h1);
        else select_net<8>(p, cnt, o, g0, &h1);
        if (o + 1 < cnt) { *g1 = h1; same_bin = true; }
    }
    if (!want1 || same_bin) return true;
    const bool atom1 = locate(v, r + 1, g1, &start, &cnt, &o);
    if (!atom1) *g1 = bin_min(v.buf + start * COLS, cnt);     // o == 0: first value of its bin (any size)
    return true;
}

// np.quantile(method='linear') on the sorted column: h = (n-1) q, NumPy _lerp
// evaluated operation by operation in float64 (same arithmetic as the other K2
// kernels' plane_value)
RS_HD float quantile_value(const ColView& v, double q, bool wide, bool* ok) {
    const int n = v.n;
    if (n == 0) return qnan();
    if (!(v.hi > v.lo)) return v.lo;
    const double dn = static_cast<double>(n);
    const double h = dmul(dn - 1.0, q);
    const double fl = floor(h);
    int l0 = static_cast<int>(fl);
    const int top = n - 2;
    l0 = l0 > top ? top : l0;
    l0 = l0 < 0 ? 0 : l0;
    const double tt = h - fl;
    float f0 = v.lo, f1 = v.lo;
    if (n >= 2) {
        if (!rank_pair(v, l0, true, wide, &f0, &f1)) { *ok = false; return qnan(); }
    }
    const double x0 = static_cast<double>(f0), x1 = static_cast<double>(f1);
    const double d = x1 - x0;
    const double rq = (tt >= 0.5) ? dsub(x1, dmul(d, 1.0 - tt)) : dadd(x0, dmul(d, tt));
    float r = static_cast<float>(rq);
    if (h >= dn - 1.0) r = v.hi;
    if (h < 0.0) r = v.lo;
    return r;
}

// quantile planes: item k of the column belongs to warp k % NW
template <int NW, int IPT>
RS_HD void phase_fin_quant(const Thr<IPT>& t, const Smem& s, const Args& a, int w, int c) {
    if (w >= a.nq) return;                       // warp-uniform
    ColView v = make_view(t, s, c);
    if (t.dead) v.n = 0;                         // no lane leaves early: the warp votes below
    for (int k = w; k < a.nq; k += NW) {
        const int pl = a.qplane[k];
        bool ok = true;
        // the populations of the bins around this quantile: if every column of the
        // warp has small ones, the 8-value network is enough
        bool wide = false;
#ifdef __CUDA_ARCH__
        {
            const int probe = v.n > 1 ? static_cast<int>(a.q[pl] * (v.n - 1)) : 0;
            int st = 0, cn = 0, oo = 0;
            float dv;
            const bool atom = v.n == 0 || !(v.hi > v.lo) || locate(v, probe, &dv, &st, &cn, &oo);
            wide = __any_sync(0xffffffffu, !atom && cn > 8);
        }
#endif
        const float r = quantile_value(v, a.q[pl], wide, &ok);
        if (!ok) s.colinfo[CI_FLAG * COLS + c] = 1u;
        s.res[pl * COLS + c] = r;
    }
}

// mode = most frequent value, ties -> smallest (defined in oracle/np_oracle.py).
// Equal values share a bin, so only pairs inside a bin are compared, and only in
// bins whose count can still reach the best multiplicity seen so far.
template <int NW, int IPT>
RS_HD void phase_fin_mode(const Thr<IPT>& t, const Smem& s, int w, int c) {
    if (t.dead || t.n == 0) return;
    const ColView v = make_view(t, s, c);
    if (w == 0) {
        unsigned long long k = mode_key(static_cast<uint32_t>(v.n_lo), v.lo);
        if (v.n_hi > 0) {
            const unsigned long long k2 = mode_key(static_cast<uint32_t>(v.n_hi), v.hi);
            k = k2 > k ? k2 : k;
        }
        atomic_max_u64(s.modekey + c, k);
    }
    if (!(v.hi > v.lo)) return;
    // a regular value wins with m > n_lo (lo is smaller) and m >= n_hi
    int thr = v.n_lo + 1 > v.n_hi ? v.n_lo + 1 : v.n_hi;
    thr = thr < 2 ? 2 : thr;
    uint32_t best_m = 0;
    float best_v = 0.f;
    for (int b = w; b < NBINS; b += NW) {
        const int end = static_cast<int>(v.cells[b * COLS]);
        const int start = b == 0 ? 0 : b == 1 ? v.n0tot : static_cast<int>(v.cells[(b - 1) * COLS]);
        const int cnt = end - start;
        const int hint = static_cast<int>(s.colinfo[CI_HINT * COLS + c]);
        if (cnt < thr || cnt < hint) continue;
        if (cnt > RS_BIGBIN) { s.colinfo[CI_FLAG * COLS + c] = 1u; continue; }
        const float* p = v.buf + start * COLS;
        uint32_t bm = 0;
        float bv = 0.f;
        for (int i = 1; i < cnt; ++i) {
            const float x = p[i * COLS];
            uint32_t m = 1;
            for (int j = 0; j < i; ++j) m += (p[j * COLS] == x) ? 1u : 0u;
            if (m > bm || (m == bm && x < bv)) { bm = m; bv = x; }
        }
        if (bm >= 2 && (bm > best_m || (bm == best_m && bv < best_v))) {
            best_m = bm;
            best_v = bv;
            atomic_max_u32(s.colinfo + CI_HINT * COLS + c, bm);
        }
    }
    if (best_m >= 2 && static_cast<int>(best_m) >= thr)
        atomic_max_u64(s.modekey + c, mode_key(best_m, best_v));
}

// ------------------------------------------------------------------ out
// moment planes: float64 finals from the float64 sums of the float32 partials
template <int NW, int IPT>
RS_HD float plane_value(const Thr<IPT>& t, const Smem& s, const Args& a, int pl, int c) {
    constexpr int S = NW / 8;
    const int kind = a.kind[pl];
    const int n = t.n;
    if (n == 0) return qnan();
    if (kind == PK_QUANT) return s.res[pl * COLS + c];
    if (kind == PK_MODE) {
        const unsigned long long k = s.modekey[c];
        return okey_inv(0xffffffffu - static_cast<uint32_t>(k));
    }
    double S1 = 0.0, S2 = 0.0, S3 = 0.0;
    for (int sub = 0; sub < S; ++sub) {
        S1 += s.part2[(1 * S + sub) * COLS + c];
        S2 += s.part2[(2 * S + sub) * COLS + c];
        S3 += s.part2[(3 * S + sub) * COLS + c];
    }
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

}  // namespace rsel
