// k1_dss14_emul.cpp -- TEST INFRASTRUCTURE: the '%12.4f\r\n' record parser of K1
// (gsimcli_b200/csrc/k1_dss14.cuh, the code the CUDA decoder inlines) compiled
// for the host, so that the CPU suite can check its layout validation, digit
// conversion and float32-rounding guard against Python's float().  Never linked
// into libgsimcli_b200.so.
//   g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC -o tests/emul/libk1dss14_emul.so tests/emul/k1_dss14_emul.cpp
#include "../../gsimcli_b200/csrc/k1_dss14.cuh"

// text: nlines records of 14 bytes.  out[i] = value, ok[i] = 1 when the fast
// parser accepted the record (0: the CUDA path would fall through to the general
// parsers).  The two bytes after the last record are read as blanks.
extern "C" void k1_dss14_emulate(const unsigned char* text, long long nlines, float* out,
                                 unsigned char* ok) {
    for (long long i = 0; i < nlines; ++i) {
        unsigned char rec[16];
        for (int k = 0; k < 16; ++k) {
            const long long o = i * 14 + k;
            rec[k] = o < nlines * 14 ? text[o] : ' ';
        }
        unsigned u[4];
        memcpy(u, rec, 16);
        float v = 0.f;
        ok[i] = k1_parse_dss14_words(u[0], u[1], u[2], u[3], &v) ? 1 : 0;
        out[i] = v;
    }
}

// the 8-records-per-thread form (k1_decode_dss14x8): groups of 112 bytes as 28
// words, record k taken from its compile-time word / half-word position.
// nlines must be a multiple of 8.
extern "C" void k1_dss14_emulate_x8(const unsigned char* text, long long nlines, float* out,
                                    unsigned char* ok) {
    for (long long g = 0; g < nlines / 8; ++g) {
        unsigned w[28];
        memcpy(w, text + g * 112, 112);
        for (int k = 0; k < 8; ++k) {
            float v = 0.f;
            ok[g * 8 + k] = k1_parse_dss14_group(w, k, &v) ? 1 : 0;
            out[g * 8 + k] = v;
        }
    }
}
