// k1_parse.cuh -- the general number parsers of K1 (Python float() semantics on a
// line of text -> float64 correctly rounded -> float32), shared with the host
// emulation in tests/emul (this header compiled by g++ and fuzzed against Python's
// float() on the CPU).  Included by k1_decode.cu inside its anonymous namespace.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define K1P_DEV __device__ __forceinline__
#define K1P_FN __device__
#define K1P_CONST __device__
#else
#define K1P_DEV static inline
#define K1P_FN static
#define K1P_CONST static
// the CUDA built-ins the parsers use, for the host
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline float __double2float_rn(double d) { return static_cast<float>(d); }
static inline int __ffs(unsigned v) { return __builtin_ffs(static_cast<int>(v)); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    return static_cast<unsigned>(((static_cast<unsigned long long>(hi) << 32) | lo) >> (sh & 31u));
}
#endif

K1P_DEV bool is_space(unsigned c) {
    return c == ' ' || (c >= 9 && c <= 13);
}

K1P_CONST const double k_pow10[23] = {
    1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
    1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// correctly rounded m * 10^e10 for 0 < m < 2^64, -21 <= e10 <= 27 (exact
// 128-bit arithmetic, round half to even).
K1P_FN double scale_exact(unsigned long long m, int e10) {
    typedef unsigned __int128 u128;
    if (e10 >= 0) {
        u128 v = m;
        for (int i = 0; i < e10; ++i) {
            if (v >> 123) return __longlong_as_double(0x7ff8000000000000LL);  // caller: EPARSE
            v *= 10;
        }
        // round v to 53 bits
        int hb = 127;
        while (!((v >> hb) & 1)) --hb;
        if (hb <= 52) return static_cast<double>(static_cast<unsigned long long>(v));
        const int sh = hb - 52;
        u128 q = v >> sh;
        const u128 rem = v & ((static_cast<u128>(1) << sh) - 1);
        const u128 half = static_cast<u128>(1) << (sh - 1);
        if (rem > half || (rem == half && (q & 1))) ++q;
        return ldexp(static_cast<double>(static_cast<unsigned long long>(q)), sh);
    }
    // m / 10^k: scale m up so the quotient keeps >= 64 significant bits
    u128 den = 1;
    for (int i = 0; i < -e10; ++i) den *= 10;       // <= 10^21 < 2^70
    // numerator m << s with s chosen so that num < 2^127 and quotient >= 2^55
    int mb = 63;
    while (!((m >> mb) & 1)) --mb;
    const int s = 126 - mb;
    const u128 num = static_cast<u128>(m) << s;
    u128 q = num / den;
    const u128 r = num % den;
    int hb = 127;
    while (!((q >> hb) & 1)) --hb;
    // q >= 2^126 / 2^70 = 2^56: more than 53 significant bits
    const int sh = hb - 52;
    u128 top = q >> sh;
    const u128 low = q & ((static_cast<u128>(1) << sh) - 1);
    const u128 half = static_cast<u128>(1) << (sh - 1);
    const bool sticky = (r != 0);
    if (low > half || (low == half && (sticky || (top & 1)))) ++top;
    return ldexp(static_cast<double>(static_cast<unsigned long long>(top)), sh - s);
}

// Python float(bytes) on text[0:len).  Returns false on a malformed line.
template <typename Ptr>
K1P_FN bool parse_float(Ptr p, int len, float* out) {
    int i = 0;
    while (i < len && is_space(p[i])) ++i;
    while (len > i && is_space(p[len - 1])) --len;
    if (i >= len) return false;
    bool negv = false;
    if (p[i] == '+' || p[i] == '-') { negv = (p[i] == '-'); ++i; }
    if (i >= len) return false;
    unsigned c = p[i] | 0x20;
    if (c == 'i' || c == 'n') {
        // inf / infinity / nan
        const int rem = len - i;
        auto eq = [&](const char* w, int n) {
            if (rem != n) return false;
            for (int k = 0; k < n; ++k)
                if ((p[i + k] | 0x20) != static_cast<unsigned>(w[k])) return false;
            return true;
        };
        if (eq("inf", 3) || eq("infinity", 8)) {
            *out = __int_as_float(negv ? 0xff800000 : 0x7f800000);
            return true;
        }
        if (eq("nan", 3)) { *out = __int_as_float(0x7fc00000); return true; }
        return false;
    }
    unsigned long long m = 0;
    int ndig = 0;        // significant digits accumulated into m
    int nd_any = 0;      // digits seen at all
    int e10 = 0;
    bool dropped_nonzero = false;
    bool seen_dot = false;
    for (; i < len; ++i) {
        const unsigned ch = p[i];
        if (ch >= '0' && ch <= '9') {
            ++nd_any;
            if (ndig < 19) {
                if (m != 0 || ch != '0') { m = m * 10 + (ch - '0'); ++ndig; }
                if (seen_dot) --e10;
            } else {
                if (ch != '0') dropped_nonzero = true;
                if (!seen_dot) ++e10;
            }
        } else if (ch == '.' && !seen_dot) {
            seen_dot = true;
        } else {
            break;
        }
    }
    if (nd_any == 0) return false;
    if (i < len) {
        if ((p[i] | 0x20) != 'e') return false;
        ++i;
        bool eneg = false;
        if (i < len && (p[i] == '+' || p[i] == '-')) { eneg = (p[i] == '-'); ++i; }
        if (i >= len) return false;
        int ex = 0;
        for (; i < len; ++i) {
            const unsigned ch = p[i];
            if (ch < '0' || ch > '9') return false;
            if (ex < 100000) ex = ex * 10 + (ch - '0');
        }
        e10 += eneg ? -ex : ex;
    }
    double d;
    if (m == 0) {
        d = 0.0;
    } else if (dropped_nonzero) {
        return false;                      // > 19 significant digits: unsupported
    } else {
        // strip trailing zeros of m only when that is what brings a padded number
        // back onto the fast path (the quotient m / 10^k is the same rational
        // either way, and the 64-bit divisions are the slowest thing here)
        while (ndig > 15 && e10 < 0 && (m % 10) == 0) { m /= 10; ++e10; --ndig; }
        if (ndig <= 15 && e10 >= -22 && e10 <= 22) {
            const double dm = static_cast<double>(static_cast<long long>(m));
            d = (e10 < 0) ? __ddiv_rn(dm, k_pow10[-e10]) : __dmul_rn(dm, k_pow10[e10]);
        } else if (e10 >= -21 && e10 <= 27) {
            d = scale_exact(m, e10);
            if (d != d) return false;
        } else {
            return false;
        }
    }
    if (negv) d = -d;
    *out = __double2float_rn(d);
    return true;
}

// ------------------------------------------------------------ fixed width
// Fast path for the record DSS writes ('%W.Pf' + EOL, W <= 16 bytes): blanks, an
// optional sign, digits with exactly one '.', then only end-of-line blanks.
// The 16 bytes are classified in registers (SWAR digit / blank masks -> one
// 16-bit mask each), the structure is validated on the masks, and the digits
// are accumulated by one short Horner loop; the conversion is the same
// correctly-rounded divide as in parse_float, so the result is bit-identical.
// Anything else (exponents, inf/nan, no '.', > 15 digits, odd characters)
// returns false and takes parse_float.
K1P_DEV unsigned k1_movemask(unsigned hibits) {
    // bit i of the result = bit 7 of byte i
    return (((hibits >> 7) & 0x01010101u) * 0x01020408u) >> 24;
}

K1P_DEV bool parse_fixed_fast(const unsigned char* p, int W, float* out) {
    // (only the low address bits matter: shared and global windows are both
    // 4-byte aligned in the generic address space)
    const unsigned addr = static_cast<unsigned>(reinterpret_cast<uintptr_t>(p));
    const unsigned sh = (addr & 3u) * 8u;
    const unsigned* wp = reinterpret_cast<const unsigned*>(p - (addr & 3u));
    unsigned w[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) w[k] = wp[k];
    unsigned nd16 = 0, sp16 = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned u = __funnelshift_r(w[k], w[k + 1], sh);
        const unsigned t = u ^ 0x30303030u;                       // digits -> 0..9
        const unsigned nd = (((t & 0x7f7f7f7fu) + 0x76767676u) | t) & 0x80808080u;
        const unsigned z = u ^ 0x20202020u;                       // ' ' -> 0
        const unsigned nz = (((z & 0x7f7f7f7fu) + 0x7f7f7f7fu) | z) & 0x80808080u;
        nd16 |= k1_movemask(nd) << (4 * k);
        sp16 |= (k1_movemask(nz) ^ 0xfu) << (4 * k);
    }
    const unsigned inw = (1u << W) - 1u;
    const unsigned dg = ~nd16 & inw;
    if (dg == 0u) return false;
    const int f = __ffs(dg) - 1;
    const int l = 31 - __clz(dg);
    const unsigned inner = nd16 & ((2u << l) - 1u) & ~((1u << f) - 1u);
    if (__popc(inner) != 1) return false;
    const int pdot = __ffs(inner) - 1;
    if (p[pdot] != '.') return false;
    for (int i = l + 1; i < W; ++i) {                   // end-of-line blanks only
        const unsigned c = p[i];
        if (!(c == ' ' || (c >= 9 && c <= 13))) return false;
    }
    int lead = f;                                       // bytes before the sign / first digit
    bool negv = false;
    if (f > 0) {
        const unsigned c = p[f - 1];
        if (c == '-' || c == '+') { negv = (c == '-'); lead = f - 1; }
    }
    const unsigned leadmask = (1u << lead) - 1u;
    if ((sp16 & leadmask) != leadmask) return false;   // something else than ' ' in front
    if (l - f > 15) return false;                       // > 15 digits: not the fast path
    const int nfrac = l - pdot;
    double d;
    if (l - f <= 9) {                                   // <= 9 digits: 32-bit Horner
        unsigned m = 0;
        for (int i = f; i < pdot; ++i) m = m * 10u + (p[i] - '0');
        for (int i = pdot + 1; i <= l; ++i) m = m * 10u + (p[i] - '0');
        d = static_cast<double>(m);
        if (m != 0u) d = __ddiv_rn(d, k_pow10[nfrac]);
    } else {
        unsigned long long m = 0;
        for (int i = f; i < pdot; ++i) m = m * 10ull + (p[i] - '0');
        for (int i = pdot + 1; i <= l; ++i) m = m * 10ull + (p[i] - '0');
        d = static_cast<double>(static_cast<long long>(m));
        if (m != 0ull) d = __ddiv_rn(d, k_pow10[nfrac]);
    }
    if (negv) d = -d;
    *out = __double2float_rn(d);
    return true;
}
