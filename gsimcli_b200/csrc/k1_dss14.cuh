// k1_dss14.cuh -- the parser of K1 for the DSS map record itself (shared with the
// host emulation in tests/emul, which compiles this header with g++ and checks it
// against Python's float() on the CPU).
//
// The record: '%12.4f\r\n' (interface/gui.py:1414)
// 14 bytes: blanks, an optional '-', the integer digits right-aligned in bytes
// 0..6, '.' at byte 7, four fraction digits, CR, LF.  Everything is done on four
// 32-bit words: SWAR byte classes (digit / blank / minus) validate the layout,
// each word of digits becomes a number with two multiply-adds (blanks count as
// leading zeros), and the decimal -> float64 division by 10^4 is a multiply by
// the rounded reciprocal, accepted only when the float64 result is provably far
// enough from a float32 rounding boundary for the final float32 to equal the
// correctly rounded path (otherwise, and for any other layout, the caller
// falls through to the general parsers).  Same value, bit for bit.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define K1_HD __device__ __forceinline__
#else
#define K1_HD static inline
#endif

K1_HD unsigned k1_notdigit(unsigned u) {      // 0x80 per non-digit byte
    const unsigned t = u ^ 0x30303030u;
    return (((t & 0x7f7f7f7fu) + 0x76767676u) | t) & 0x80808080u;
}
K1_HD unsigned k1_nonzero(unsigned z) {       // 0x80 per non-zero byte
    return (((z & 0x7f7f7f7fu) + 0x7f7f7f7fu) | z) & 0x80808080u;
}
#ifdef __CUDA_ARCH__
// the exact division is needed for about one record in 10^7; kept out of line so
// that the compiler cannot if-convert it into the common path (inlined, its
// reciprocal iteration -- MUFU.RCP64H + 7 DFMA/DMUL + fix-up tests -- ran for
// every record and the result was thrown away by a select)
__device__ __noinline__ double k1_div1e4_exact(double m) { return __ddiv_rn(m, 1.0e4); }
#endif
K1_HD unsigned k1_digits4(unsigned t) {       // 4 digit bytes (byte 0 first) -> number
    t = (t * 10u + (t >> 8)) & 0x00ff00ffu;
    return (t * 100u + (t >> 16)) & 0xffffu;
}

// u0..u3 = bytes 0..15 of the record as little-endian words.  Returns false for
// any layout other than blanks* '-'? digits+ '.' dddd CR LF (the caller falls
// through to the general parsers).
K1_HD bool k1_parse_dss14_words(unsigned u0, unsigned u1, unsigned u2, unsigned u3, float* out) {
    // fixed characters
    bool ok = ((u1 >> 24) == 0x2eu) && ((u3 & 0xffffu) == 0x0a0du);
    // byte classes of the 7 integer-part bytes
    const unsigned nd0 = k1_notdigit(u0), nd1 = k1_notdigit(u1);
    const unsigned D0 = nd0 ^ 0x80808080u, D1 = (nd1 ^ 0x80808080u) & 0x00808080u;
    const unsigned M0 = k1_nonzero(u0 ^ 0x2d2d2d2du) ^ 0x80808080u;
    const unsigned M1 = (k1_nonzero(u1 ^ 0x2d2d2d2du) ^ 0x80808080u) & 0x00808080u;
    const unsigned S0 = k1_nonzero(u0 ^ 0x20202020u) ^ 0x80808080u;
    const unsigned S1 = (k1_nonzero(u1 ^ 0x20202020u) ^ 0x80808080u) & 0x00808080u;
    ok = ok && ((D0 | M0 | S0) == 0x80808080u) && ((D1 | M1 | S1) == 0x00808080u);
    ok = ok && (D1 & 0x00800000u);                        // byte 6 is a digit
    // digits form a suffix of bytes 0..6 (no non-digit after a digit) and a '-'
    // is followed by a digit: with both, the layout is blanks* '-'? digits+
    const unsigned long long D = (static_cast<unsigned long long>(D1) << 32) | D0;
    const unsigned long long M = (static_cast<unsigned long long>(M1) << 32) | M0;
    const unsigned long long in7 = 0x0080808080808080ull;
    ok = ok && ((((D << 8) & ~D) & in7) == 0ull) && ((((M << 8) & ~D) & (in7 | (1ull << 63))) == 0ull);
    ok = ok && (k1_notdigit(u2) == 0u);
    if (!ok) return false;
    // digit bytes keep their low nibble, everything else (blank, '-') is zero
    const unsigned v0 = k1_digits4(u0 & ((D0 >> 7) * 0x0fu));
    const unsigned v1 = k1_digits4(u1 & ((D1 >> 7) * 0x0fu));    // d4 d5 d6 0 = 10 * (d4 d5 d6)
    const unsigned v2 = k1_digits4(u2 & 0x0f0f0f0fu);
    const unsigned ip10 = v0 * 10000u + v1;                      // 10 * integer part < 10^8
    // m = integer * 10^4 + fraction, exact in float64 (< 10^11)
    const double m = fma(static_cast<double>(ip10), 1000.0, static_cast<double>(v2));
#ifdef __CUDA_ARCH__
    double x = __dmul_rn(m, 1.0e-4);
#else
    double x = m * 1.0e-4;
#endif
    // x is within 2^-51 (relative) of the correctly rounded m / 10^4.  The
    // float32 rounding boundaries of a float64 are the values whose low 29
    // mantissa bits are 0x10000000: stay 16 units away from them (and from the
    // wrap of the bits below, 0 and 2^29) or take the exact division.
    unsigned long long xb;
    memcpy(&xb, &x, 8);
    const unsigned low = static_cast<unsigned>(xb) & 0x1fffffffu;
    if (((low + 16u) & 0x0fffffffu) < 32u) {
#ifdef __CUDA_ARCH__
        x = k1_div1e4_exact(m);
#else
        x = m / 1.0e4;
#endif
    }
    if (M != 0ull) x = -x;
    *out = static_cast<float>(x);
    return true;
}

// Record k (0..7) of a group of 8 records = 112 bytes held as 28 little-endian
// words (k1_decode_dss14x8): even records start on a word, odd ones 16 bits in.
K1_HD bool k1_parse_dss14_group(const unsigned (&w)[28], int k, float* out) {
    const int b = (14 * k) >> 2;
    unsigned u0, u1, u2, u3;
    if ((k & 1) == 0) {
        u0 = w[b]; u1 = w[b + 1]; u2 = w[b + 2]; u3 = w[b + 3];
    } else {
        u0 = (w[b] >> 16) | (w[b + 1] << 16);
        u1 = (w[b + 1] >> 16) | (w[b + 2] << 16);
        u2 = (w[b + 2] >> 16) | (w[b + 3] << 16);
        u3 = w[b + 3] >> 16;
    }
    return k1_parse_dss14_words(u0, u1, u2, u3, out);
}
