// k1_parse_emul.cpp -- TEST INFRASTRUCTURE: the general number parsers of K1
// (gsimcli_b200/csrc/k1_parse.cuh, the code the CUDA decoder compiles) built for
// the host, so that the CPU suite can fuzz them against Python's float().  Never
// linked into libgsimcli_b200.so.
//   g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC -o tests/emul/libk1parse_emul.so tests/emul/k1_parse_emul.cpp
#include "../../gsimcli_b200/csrc/k1_parse.cuh"

// text: nbytes of lines; line i = [off[2 i], off[2 i + 1]) (without its '\n').  The
// buffer must be readable 32 bytes past nbytes (the register fast path loads
// aligned words).  which: 0 = as k1_decode_general chooses (fast path for short
// lines, else / on refusal the full parser), 1 = parse_float only, 2 = the fast
// path only.  ok[i] = 1 when the line was accepted, out[i] = its float32.
extern "C" void k1_parse_emulate(const unsigned char* text, long long nbytes, const long long* off,
                                 long long nlines, int which, float* out, unsigned char* ok) {
    for (long long i = 0; i < nlines; ++i) {
        const unsigned char* lp = text + off[2 * i];
        const int len = static_cast<int>(off[2 * i + 1] - off[2 * i]);
        float val = 0.f;
        bool good;
        const bool fast_ok = len >= 1 && len <= 16 && off[2 * i] + 24 <= nbytes;
        if (which == 1) good = parse_float(lp, len, &val);
        else if (which == 2) good = fast_ok && parse_fixed_fast(lp, len, &val);
        else good = (fast_ok && parse_fixed_fast(lp, len, &val)) || parse_float(lp, len, &val);
        ok[i] = good ? 1 : 0;
        out[i] = val;
    }
}
