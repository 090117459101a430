// k1_decode.cu -- K1: GSLIB ASCII grid decoder (text -> float32), and its inverse
//
// Replaces the reference's per-line `grid.readline()` + `float()`
// (GSIMCLI/tools/grid.py:673, 820-823; tools/utils.py:66-78) and
// np.loadtxt(skiprows) (grid.py:326).  One number per line, Python float()
// syntax.  Conversion is decimal -> float64 correctly rounded (what float()
// returns) -> float32 round-to-nearest-even (our storage cast):
//   * <= 15 significant digits and |e10| <= 22: Clinger's fast path, one exact
//     u64->f64 conversion and one correctly rounded multiply/divide;
//   * up to 19 digits, -21 <= e10 <= 27: exact 128-bit integer arithmetic with
//     round-half-even;
//   * anything else: GSB_EPARSE (never a host fallback).
// Two line-addressing modes:
//   * fixed-width records (line_width > 0): O(1) offsets, a CTA stages a
//     contiguous run of records through shared memory with 16-byte loads;
//   * general: a newline index is built on the device (16 bytes per thread,
//     byte-compare masks, block scan of popcounts), then one thread per line.
#include "gsb_common.cuh"
#include "k1_dss14.cuh"

namespace {

constexpr int K1_THREADS = 256;
constexpr int K1_LINES_PER_THREAD = 2;
constexpr int K1_LPB = K1_THREADS * K1_LINES_PER_THREAD;

#include "k1_parse.cuh"

// ---- the DSS map record itself: '%12.4f\r\n' (interface/gui.py:1414): k1_dss14.cuh
__device__ __forceinline__ bool parse_dss14(const unsigned char* p, float* out) {
    const unsigned addr = static_cast<unsigned>(reinterpret_cast<uintptr_t>(p));
    const unsigned sh = (addr & 3u) * 8u;
    const unsigned* wp = reinterpret_cast<const unsigned*>(p - (addr & 3u));
    const unsigned a0 = wp[0], a1 = wp[1], a2 = wp[2], a3 = wp[3], a4 = wp[4];
    return k1_parse_dss14_words(__funnelshift_r(a0, a1, sh),      // bytes 0..3   integer part
                                __funnelshift_r(a1, a2, sh),      // bytes 4..7   integer part, '.'
                                __funnelshift_r(a2, a3, sh),      // bytes 8..11  fraction
                                __funnelshift_r(a3, a4, sh),      // bytes 12..13 CR LF
                                out);
}

__global__ void __launch_bounds__(K1_THREADS)
k1_decode_fixed(const char* __restrict__ text, long long nbytes, int W, long long nlines,
                float* __restrict__ out, long long* bad_line) {
    extern __shared__ __align__(16) unsigned char sm[];
    const long long line0 = static_cast<long long>(blockIdx.x) * K1_LPB;
    const long long nl = min(static_cast<long long>(K1_LPB), nlines - line0);
    if (nl <= 0) return;
    const long long b0 = line0 * W;
    const long long b1 = b0 + nl * W;
    // stage [a0, a1) with 16-byte loads (a0 = b0 rounded down to 16)
    const uintptr_t base = reinterpret_cast<uintptr_t>(text);
    const long long mis = static_cast<long long>((base + b0) & 15);
    const long long a0 = b0 - mis;
    const long long nvec = (b1 - a0 + 15) / 16;
    for (long long v = threadIdx.x; v < nvec; v += K1_THREADS) {
        const long long off = a0 + v * 16;
        int4 w;
        if (off + 16 <= nbytes && off >= 0) {
            w = __ldg(reinterpret_cast<const int4*>(text + off));
        } else {
            unsigned char tmp[16];
            for (int k = 0; k < 16; ++k) {
                const long long o = off + k;
                tmp[k] = (o >= 0 && o < nbytes) ? static_cast<unsigned char>(text[o]) : ' ';
            }
            w = *reinterpret_cast<int4*>(tmp);
        }
        reinterpret_cast<int4*>(sm)[v] = w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K1_LINES_PER_THREAD; ++k) {
        const int li = threadIdx.x + k * K1_THREADS;
        if (li < nl) {
            const unsigned char* p = sm + mis + static_cast<long long>(li) * W;
            float val;
            bool ok = (W == 14 && parse_dss14(p, &val)) ||
                      ((p[W - 1] == '\n') &&
                       ((W <= 16 && parse_fixed_fast(p, W, &val)) || parse_float(p, W, &val)));
            if (ok) out[line0 + li] = val;
            else atomicMin(reinterpret_cast<unsigned long long*>(bad_line),
                           static_cast<unsigned long long>(line0 + li));
        }
    }
}

// ------------------------------------------------------------ DSS records, 8 per thread
// EXPERIMENTAL (stack option "k1_algo" = 1, never the default: written after the
// round's GPU budget was spent; bench.py times it in its `experimental` leg and
// compares the decoded values bit for bit with the default kernel's).
// '%12.4f\r\n' records are 14 bytes, so 8 of them are 112 bytes = seven 16-byte
// vectors: a thread loads them straight into registers (no shared-memory staging,
// no per-line word loads), every record starts at a compile-time word / half-word
// position (even records are word aligned, odd ones 16 bits in), and the 8 results
// leave as two 16-byte stores.  Same parser (k1_dss14.cuh), same fall-back parsers
// for any record it refuses.  Needs the first record 16-byte aligned.
// any spelling the record parser refuses: the general parsers, from global memory
// (out of line: it is inlined nowhere near the common path)
__device__ __noinline__ void k1_dss14_slow(const char* text, long long nbytes, long long li, float* out,
                                           long long* bad_line) {
    const long long off = li * 14;
    const unsigned char* lp = reinterpret_cast<const unsigned char*>(text) + off;
    float val;
    const bool ok = (lp[13] == '\n') &&
                    ((off + 24 <= nbytes && parse_fixed_fast(lp, 14, &val)) || parse_float(lp, 14, &val));
    if (ok) out[li] = val;
    else atomicMin(reinterpret_cast<unsigned long long*>(bad_line), static_cast<unsigned long long>(li));
}

template <bool VEC_STORE>
__global__ void __launch_bounds__(K1_THREADS)
k1_decode_dss14x8(const char* __restrict__ text, long long nbytes, long long nlines,
                  float* __restrict__ out, long long* bad_line) {
    const long long g = static_cast<long long>(blockIdx.x) * K1_THREADS + threadIdx.x;
    const long long ngroups = nlines >> 3;
    if (g > ngroups) return;
    auto slow = [&](long long li) { k1_dss14_slow(text, nbytes, li, out, bad_line); };
    if (g == ngroups) {                 // the last nlines % 8 records
        for (long long li = ngroups * 8; li < nlines; ++li) slow(li);
        return;
    }
    const uint4* p = reinterpret_cast<const uint4*>(text + g * 112);
    unsigned w[28];
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const uint4 v = __ldg(p + k);
        w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
    }
    float f[8];
    unsigned okmask = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k1_parse_dss14_group(w, k, &f[k])) okmask |= 1u << k;   // k is compile-time after unrolling
    }
    float* o = out + g * 8;
    if (okmask == 0xffu) {
        if (VEC_STORE) {
            reinterpret_cast<float4*>(o)[0] = make_float4(f[0], f[1], f[2], f[3]);
            reinterpret_cast<float4*>(o)[1] = make_float4(f[4], f[5], f[6], f[7]);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = f[k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (okmask & (1u << k)) o[k] = f[k];
            else slow(g * 8 + k);
        }
    }
}

// ------------------------------------------------------------ general form
// pass 1: newline count per block (each thread 16 bytes)
__global__ void __launch_bounds__(K1_THREADS)
k1_count_newlines(const char* __restrict__ text, long long nbytes, unsigned* block_counts) {
    __shared__ unsigned wsum[K1_THREADS / 32];
    const long long off = (static_cast<long long>(blockIdx.x) * K1_THREADS + threadIdx.x) * 16;
    unsigned cnt = 0;
    if (off + 16 <= nbytes && ((reinterpret_cast<uintptr_t>(text) & 15) == 0)) {
        const int4 w = __ldg(reinterpret_cast<const int4*>(text + off));
        const unsigned ws[4] = {(unsigned)w.x, (unsigned)w.y, (unsigned)w.z, (unsigned)w.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // bytes equal to '\n' (0x0a): SWAR zero-byte test on x ^ 0x0a0a0a0a
            const unsigned x = ws[k] ^ 0x0a0a0a0au;
            const unsigned t = ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu);
            cnt += __popc(t);
        }
    } else {
        for (int k = 0; k < 16; ++k)
            if (off + k < nbytes && text[off + k] == '\n') ++cnt;
    }
    const unsigned wtot = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = wtot;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int w = 0; w < K1_THREADS / 32; ++w) t += wsum[w];
        block_counts[blockIdx.x] = t;
    }
}

// pass 2: exclusive scan of block counts (single CTA, sequential chunks)
__global__ void __launch_bounds__(1024) k1_scan_blocks(unsigned* counts, long long nblocks,
                                                       unsigned long long* total) {
    __shared__ unsigned long long carry;
    __shared__ unsigned wsum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (long long base = 0; base < nblocks; base += 1024) {
        const long long i = base + threadIdx.x;
        const unsigned v = (i < nblocks) ? counts[i] : 0;
        unsigned x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned w = wsum[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned y = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += y;
            }
            wsum[threadIdx.x] = w;
        }
        __syncthreads();
        const unsigned wprev = (threadIdx.x >> 5) ? wsum[(threadIdx.x >> 5) - 1] : 0;
        const unsigned long long excl = carry + wprev + x - v;
        // counts[] becomes the exclusive prefix; newline indices stay < 2^32
        if (i < nblocks) counts[i] = static_cast<unsigned>(excl);
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// pass 3: write the byte offset of every newline, line_end[rank] = position
__global__ void __launch_bounds__(K1_THREADS)
k1_index_newlines(const char* __restrict__ text, long long nbytes,
                  const unsigned* __restrict__ block_prefix, unsigned* line_end,
                  long long max_lines) {
    __shared__ unsigned wsum[K1_THREADS / 32];
    const long long off = (static_cast<long long>(blockIdx.x) * K1_THREADS + threadIdx.x) * 16;
    unsigned mask = 0;   // bit k: byte off+k is '\n'
    for (int k = 0; k < 16; ++k)
        if (off + k < nbytes && text[off + k] == '\n') mask |= 1u << k;
    const unsigned cnt = __popc(mask);
    unsigned x = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
    __syncthreads();
    unsigned wprev = 0;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) wprev += wsum[w];
    unsigned rank = block_prefix[blockIdx.x] + wprev + x - cnt;
    while (mask) {
        const int k = __ffs(mask) - 1;
        mask &= mask - 1;
        if (rank < max_lines) line_end[rank] = static_cast<unsigned>(off + k);
        ++rank;
    }
}

// pass 4: one thread per requested line
__global__ void __launch_bounds__(K1_THREADS)
k1_decode_general(const char* __restrict__ text, long long nbytes,
                  const unsigned* __restrict__ line_end, long long n_newlines,
                  long long first_file_line, long long nlines, float* __restrict__ out,
                  long long* bad_line) {
    const long long i = static_cast<long long>(blockIdx.x) * K1_THREADS + threadIdx.x;
    if (i >= nlines) return;
    const long long fl = first_file_line + i;     // line index in the file
    const long long start = fl == 0 ? 0 : static_cast<long long>(line_end[fl - 1]) + 1;
    const long long end = fl < n_newlines ? static_cast<long long>(line_end[fl]) : nbytes;
    float val;
    const unsigned char* lp = reinterpret_cast<const unsigned char*>(text) + start;
    const int len = static_cast<int>(end - start);
    // short lines take the register fast path too; its 5 aligned word loads may
    // reach up to 7 bytes past the line, so stay clear of the end of the text
    bool ok = (start <= end) && (end - start < 4096) &&
              ((len >= 1 && len <= 16 && start + 24 <= nbytes && parse_fixed_fast(lp, len, &val)) ||
               parse_float(lp, len, &val));
    if (ok) out[i] = val;
    else atomicMin(reinterpret_cast<unsigned long long*>(bad_line),
                   static_cast<unsigned long long>(i));
}

// ------------------------------------------------------------------ encoder
// fixed-width '%<width>.<prec>f' + EOL, correctly rounded for values whose
// scaled magnitude fits 2^53 (all GSLIB-style data); one thread per value
__global__ void __launch_bounds__(K1_THREADS)
k1_encode_fixed(const float* __restrict__ row, long long nn, int width, int prec, int eolmode,
                char* __restrict__ text) {
    const long long i = static_cast<long long>(blockIdx.x) * K1_THREADS + threadIdx.x;
    if (i >= nn) return;
    const bool crlf = (eolmode & 1) != 0, left = (eolmode & 2) != 0;
    const int W = width + (crlf ? 2 : 1);
    char* p = text + i * W;
    const double v = static_cast<double>(row[i]);
    const bool negv = __double_as_longlong(v) < 0;       // sign bit: printf prints -0.000000
    const double av = fabs(v);
    char tmp[40];
    int n = 0;
    if (v != v) {                       // printf: "nan"
        tmp[n++] = 'n'; tmp[n++] = 'a'; tmp[n++] = 'n';
    } else if (av > 1.0e300) {          // "inf" / "-inf" (digits are stored reversed)
        tmp[n++] = 'f'; tmp[n++] = 'n'; tmp[n++] = 'i';
        if (negv) tmp[n++] = '-';
    } else {
        double scale = 1.0;
        for (int k = 0; k < prec; ++k) scale *= 10.0;
        // exact decimal expansion of a binary float is finite: round-half-even on
        // the exact product (av*scale is exact: 24 mantissa bits times 10^prec,
        // prec <= 8, fits 53 bits)
        const double scaled = av * scale;
        unsigned long long q = static_cast<unsigned long long>(rint(scaled));
        const bool neg = negv;          // printf keeps the sign of values that round to zero
        for (int k = 0; k < prec; ++k) { tmp[n++] = '0' + static_cast<int>(q % 10); q /= 10; }
        if (prec > 0) tmp[n++] = '.';
        do { tmp[n++] = '0' + static_cast<int>(q % 10); q /= 10; } while (q && n < 38);
        if (neg) tmp[n++] = '-';
    }
    int pos = 0;
    if (!left)
        for (; pos < width - n; ++pos) p[pos] = ' ';
    for (int k = n - 1; k >= 0 && pos < width; --k) p[pos++] = tmp[k];
    for (; pos < width; ++pos) p[pos] = ' ';
    if (crlf) { p[width] = '\r'; p[width + 1] = '\n'; }
    else p[width] = '\n';
}

}  // namespace

int gsb_launch_decode(gsb_stack* s, int64_t r, const char* dtext, size_t nbytes,
                      int header_lines, int line_width, int64_t first_line, int64_t n0,
                      int64_t nlines, int64_t* d_bad_line, cudaStream_t st) {
    if (nlines <= 0) return GSB_OK;
    float* out = s->data + r * s->pitch + n0;
    if (line_width > 0) {
        GSB_REQUIRE(header_lines == 0,
                    "fixed-width decode expects the header to be stripped by the caller");
        GSB_REQUIRE(static_cast<size_t>(first_line + nlines) * line_width <= nbytes,
                    "text too short: %lld lines of %d bytes need %lld bytes, have %zu",
                    (long long)(first_line + nlines), line_width,
                    (long long)(first_line + nlines) * line_width, nbytes);
        dtext += first_line * line_width;
        nbytes -= static_cast<size_t>(first_line) * line_width;
        GSB_REQUIRE(line_width >= 2 && line_width <= 64, "line_width %d out of range",
                    line_width);
        if (s->k1_algo == 1 && line_width == 14 && (reinterpret_cast<uintptr_t>(dtext) & 15) == 0) {
            // experimental: 8 records per thread (k1_decode_dss14x8)
            const long long threads = (nlines >> 3) + 1;
            const int blocks8 = static_cast<int>((threads + K1_THREADS - 1) / K1_THREADS);
            if ((reinterpret_cast<uintptr_t>(out) & 15) == 0)
                k1_decode_dss14x8<true><<<blocks8, K1_THREADS, 0, st>>>(
                    dtext, static_cast<long long>(nbytes), nlines, out, reinterpret_cast<long long*>(d_bad_line));
            else
                k1_decode_dss14x8<false><<<blocks8, K1_THREADS, 0, st>>>(
                    dtext, static_cast<long long>(nbytes), nlines, out, reinterpret_cast<long long*>(d_bad_line));
            GSB_CUDA(cudaGetLastError());
            return GSB_OK;
        }
        const long long blocks = (nlines + K1_LPB - 1) / K1_LPB;
        const size_t smem = static_cast<size_t>(K1_LPB) * line_width + 48;
        GSB_CUDA(cudaFuncSetAttribute(k1_decode_fixed,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      K1_LPB * 64 + 48));
        k1_decode_fixed<<<static_cast<int>(blocks), K1_THREADS, smem, st>>>(
            dtext, static_cast<long long>(nbytes), line_width, nlines, out,
            reinterpret_cast<long long*>(d_bad_line));
        GSB_CUDA(cudaGetLastError());
        return GSB_OK;
    }
    // general form: newline index in scratch (after the 64 MB host-staging area)
    GSB_REQUIRE(nbytes < (1ull << 32), "general decode handles files < 4 GiB");
    const long long nblk = (static_cast<long long>(nbytes) + K1_THREADS * 16 - 1) / (K1_THREADS * 16);
    void* g = nullptr;
    const size_t need = static_cast<size_t>(nblk) * 4 + 16 + static_cast<size_t>(nbytes / 2 + 2) * 4;
    int rc = gsb_scratch(s, need, &g);
    if (rc) return rc;
    unsigned* counts = static_cast<unsigned*>(g);
    unsigned long long* total = reinterpret_cast<unsigned long long*>(
        static_cast<char*>(g) + ((static_cast<size_t>(nblk) * 4 + 15) & ~size_t(15)));
    unsigned* line_end = reinterpret_cast<unsigned*>(total + 2);
    const long long cap_lines = static_cast<long long>(nbytes / 2 + 2);
    k1_count_newlines<<<static_cast<int>(nblk), K1_THREADS, 0, st>>>(
        dtext, static_cast<long long>(nbytes), counts);
    k1_scan_blocks<<<1, 1024, 0, st>>>(counts, nblk, total);
    k1_index_newlines<<<static_cast<int>(nblk), K1_THREADS, 0, st>>>(
        dtext, static_cast<long long>(nbytes), counts, line_end, cap_lines);
    GSB_CUDA(cudaGetLastError());
    unsigned long long h_total = 0;
    GSB_CUDA(cudaMemcpyAsync(&h_total, total, 8, cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    long long n_newlines = static_cast<long long>(h_total);
    if (n_newlines > cap_lines) n_newlines = cap_lines;   // only blank lines can exceed
    // lines available = newlines (+1 if the file does not end with one)
    const long long first = header_lines + first_line;
    GSB_REQUIRE(first + nlines <= n_newlines + 1,
                "text has %lld lines, need %lld", n_newlines + 1, first + nlines);
    const long long blocks = (nlines + K1_THREADS - 1) / K1_THREADS;
    k1_decode_general<<<static_cast<int>(blocks), K1_THREADS, 0, st>>>(
        dtext, static_cast<long long>(nbytes), line_end, n_newlines, first, nlines, out,
        reinterpret_cast<long long*>(d_bad_line));
    GSB_CUDA(cudaGetLastError());
    return GSB_OK;
}

int gsb_launch_encode(const gsb_stack* s, int64_t r, int64_t n0, int64_t nn, int width,
                      int prec, int crlf, char* dtext, cudaStream_t st) {
    if (nn <= 0) return GSB_OK;
    GSB_REQUIRE(width >= 4 && width <= 32 && prec >= 0 && prec <= 9,
                "encode: width/prec out of range");
    const long long blocks = (nn + K1_THREADS - 1) / K1_THREADS;
    k1_encode_fixed<<<static_cast<int>(blocks), K1_THREADS, 0, st>>>(
        s->data + r * s->pitch + n0, nn, width, prec, crlf, dtext);
    GSB_CUDA(cudaGetLastError());
    return GSB_OK;
}
