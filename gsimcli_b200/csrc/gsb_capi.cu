// gsb_capi.cu -- the C ABI of libgsimcli_b200 (include/gsimcli_b200.h)
//
// Thin host layer: stack lifecycle, host<->device staging, plane lists and
// dispatch to the kernels in k1_decode.cu / k2_*.cu / gsb_synth.cu.  No torch
// types, no global mutable state besides the thread-local error buffer.
#include "gsb_common.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <new>
#include <mutex>

static thread_local char g_err[512] = "";

void gsb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static int grow(void** p, size_t* have, size_t need, bool pinned) {
    if (need <= *have) return GSB_OK;
    if (*p) {
        if (pinned) cudaFreeHost(*p); else cudaFree(*p);
        *p = nullptr;
        *have = 0;
    }
    // round up to limit reallocation churn
    size_t cap = (need + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
    cudaError_t e = pinned ? cudaMallocHost(p, cap) : cudaMalloc(p, cap);
    if (e != cudaSuccess) {
        gsb_set_error("allocation of %zu bytes failed: %s", cap, cudaGetErrorString(e));
        return GSB_ENOMEM;
    }
    *have = cap;
    return GSB_OK;
}

int gsb_scratch(gsb_stack* s, size_t bytes, void** out) {
    int rc = grow(&s->scratch, &s->scratch_bytes, bytes, false);
    *out = s->scratch;
    return rc;
}
int gsb_stage(gsb_stack* s, size_t bytes, void** out) {
    int rc = grow(&s->stage, &s->stage_bytes, bytes, false);
    *out = s->stage;
    return rc;
}
int gsb_pinned(gsb_stack* s, size_t bytes, void** out) {
    int rc = grow(&s->pinned, &s->pinned_bytes, bytes, true);
    *out = s->pinned;
    return rc;
}

int gsb_build_planes(uint32_t flags, double p, const double* q, int nq, GsbPlanes* out) {
    GsbPlanes pl;
    memset(&pl, 0, sizeof(pl));
    auto add = [&](int kind, double qq) {
        if (pl.n < GSB_MAX_PLANES) { pl.kind[pl.n] = (unsigned char)kind; pl.q[pl.n] = qq; }
        ++pl.n;
    };
    if (flags & GSB_MEAN) add(GSB_PK_MEAN, 0);
    if (flags & GSB_MEDIAN) add(GSB_PK_QUANT, 0.5);
    if (flags & GSB_SKEW) add(GSB_PK_SKEW, 0);
    if (flags & GSB_VAR) add(GSB_PK_VAR, 0);
    if (flags & GSB_STD) add(GSB_PK_STD, 0);
    if (flags & GSB_COEFVAR) add(GSB_PK_COEFVAR, 0);
    if (flags & GSB_PERC) {
        // grid.py:700-701 -- the operands exactly as Python evaluates them
        add(GSB_PK_QUANT, (1 - p) / 2);
        add(GSB_PK_QUANT, 1 - (1 - p) / 2);
    }
    for (int i = 0; i < nq; ++i) add(GSB_PK_QUANT, q[i]);
    if (flags & GSB_MODE) add(GSB_PK_MODE, 0);
    GSB_REQUIRE(pl.n <= GSB_MAX_PLANES, "too many output planes (%d > %d)", pl.n,
                GSB_MAX_PLANES);
    for (int i = 0; i < pl.n; ++i) {
        if (pl.kind[i] == GSB_PK_QUANT || pl.kind[i] == GSB_PK_MODE) pl.need_sort = 1;
        if (pl.kind[i] == GSB_PK_MODE) pl.need_mode = 1;
        if (pl.kind[i] == GSB_PK_QUANT)
            GSB_REQUIRE(pl.q[i] >= 0.0 && pl.q[i] <= 1.0,
                        "quantile %g outside [0, 1]", pl.q[i]);
    }
    *out = pl;
    return GSB_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point, so the
// library links against cudart only and loads on a machine without a GPU.
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static int make_tmap(gsb_stack* s) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        gsb_set_error("cuTensorMapEncodeTiled entry point unavailable");
        return GSB_ECUDA;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)s->N, (cuuint64_t)s->R};
    const cuuint64_t strides[1] = {(cuuint64_t)s->pitch * 4};
    const cuuint32_t box[2] = {32, (cuuint32_t)gsb_k2_box_rows(s->R)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = ((encode_tiled_fn)fn)(&s->tmap_k2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, s->data,
                                       dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_128B,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA);
    if (r != CUDA_SUCCESS) {
        gsb_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return GSB_ECUDA;
    }
    s->tmap_ok = 1;
    return GSB_OK;
}

extern "C" {

int gsb_version(void) { return 100; }

int gsb_last_error(char* buf, size_t n) {
    if (!buf || !n) return GSB_EINVAL;
    strncpy(buf, g_err, n - 1);
    buf[n - 1] = 0;
    return GSB_OK;
}

int gsb_device_count(int* count) {
    GSB_REQUIRE(count, "count is NULL");
    *count = 0;
    GSB_CUDA(cudaGetDeviceCount(count));
    return GSB_OK;
}

int gsb_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem,
                    char* name, size_t name_len) {
    cudaDeviceProp prop;
    GSB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (total_mem) *total_mem = prop.totalGlobalMem;
    if (name && name_len) { strncpy(name, prop.name, name_len - 1); name[name_len - 1] = 0; }
    return GSB_OK;
}

int gsb_stack_create(int device, int64_t R, int64_t N, gsb_stack** out) {
    GSB_REQUIRE(out, "out is NULL");
    *out = nullptr;
    GSB_REQUIRE(R >= 1 && N >= 1, "stack needs R >= 1 and N >= 1 (got %lld, %lld)",
                (long long)R, (long long)N);
    GSB_REQUIRE(N < (int64_t(1) << 31) - 64, "N too large for 32-bit TMA coordinates");
    GSB_CUDA(cudaSetDevice(device));
    gsb_stack* s = new (std::nothrow) gsb_stack();
    if (!s) { gsb_set_error("out of host memory"); return GSB_ENOMEM; }
    memset(s, 0, sizeof(*s));
    s->device = device;
    s->R = R;
    s->N = N;
    s->pitch = (N + 31) & ~int64_t(31);
    cudaError_t e = cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) { delete s; GSB_CUDA(e); }
    e = cudaMalloc(&s->data, (size_t)R * (size_t)s->pitch * 4);
    if (e != cudaSuccess) {
        gsb_set_error("cannot allocate a %lld x %lld float32 stack: %s", (long long)R,
                      (long long)s->pitch, cudaGetErrorString(e));
        delete s;
        return GSB_ENOMEM;
    }
    e = cudaMalloc(&s->k2_flags, (size_t)(s->pitch / 32 + 4) * sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(s->k2_flags, 0, sizeof(int));
    if (e != cudaSuccess) {
        gsb_set_error("cannot allocate the K2 tile list: %s", cudaGetErrorString(e));
        cudaFree(s->data);
        delete s;
        return GSB_ENOMEM;
    }
    cudaEventCreate(&s->ev0);
    cudaEventCreate(&s->ev1);
    int rc = make_tmap(s);
    if (rc != GSB_OK) { gsb_stack_destroy(s); return rc; }
    *out = s;
    return GSB_OK;
}

int gsb_stack_destroy(gsb_stack* s) {
    if (!s) return GSB_OK;
    cudaSetDevice(s->device);
    if (s->data) cudaFree(s->data);
    if (s->k2_flags) cudaFree(s->k2_flags);
    if (s->scratch) cudaFree(s->scratch);
    if (s->stage) cudaFree(s->stage);
    if (s->pinned) cudaFreeHost(s->pinned);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    delete s;
    return GSB_OK;
}

int gsb_stack_info(const gsb_stack* s, int64_t* R, int64_t* N, int64_t* pitch, int* device) {
    GSB_REQUIRE(s, "stack is NULL");
    if (R) *R = s->R;
    if (N) *N = s->N;
    if (pitch) *pitch = s->pitch;
    if (device) *device = s->device;
    return GSB_OK;
}

int gsb_stack_set_option(gsb_stack* s, const char* key, int64_t value) {
    GSB_REQUIRE(s && key, "NULL argument");
    if (strcmp(key, "k2_sm_reserve") == 0) {
        GSB_REQUIRE(value >= 0 && value < s->sm_count, "k2_sm_reserve out of range");
        s->k2_sm_reserve = (int)value;
        return GSB_OK;
    }
    if (strcmp(key, "k2_algo") == 0) {
        GSB_REQUIRE(value >= 0 && value <= 4, "k2_algo must be 0..4");
        s->k2_algo = (int)value;
        return GSB_OK;
    }
    if (strcmp(key, "k1_algo") == 0) {
        GSB_REQUIRE(value == 0 || value == 1, "k1_algo must be 0 or 1");
        s->k1_algo = (int)value;
        return GSB_OK;
    }
    if (strcmp(key, "k2_nw") == 0) {
        GSB_REQUIRE(value == 0 || value == 16 || value == 32, "k2_nw must be 0, 16 or 32");
        s->k2_nw = (int)value;
        return GSB_OK;
    }
    if (strcmp(key, "k2_stats") == 0) {
        s->k2_stats = value != 0;
        return GSB_OK;
    }
    gsb_set_error("unknown option '%s'", key);
    return GSB_EINVAL;
}

int gsb_debug_k2_flagged(gsb_stack* s, int* count) {
    GSB_REQUIRE(s && count, "NULL argument");
    *count = 0;
    GSB_CUDA(cudaSetDevice(s->device));
    GSB_CUDA(cudaMemcpy(count, s->k2_flags, sizeof(int), cudaMemcpyDeviceToHost));
    return GSB_OK;
}

int gsb_stack_device_ptr(const gsb_stack* s, void** dptr) {
    GSB_REQUIRE(s && dptr, "NULL argument");
    *dptr = s->data;
    return GSB_OK;
}

static int check_range(const gsb_stack* s, int64_t r0, int64_t nr, int64_t n0, int64_t nn) {
    GSB_REQUIRE(s, "stack is NULL");
    GSB_REQUIRE(r0 >= 0 && nr >= 0 && r0 + nr <= s->R, "rows [%lld, %lld) outside 0..%lld",
                (long long)r0, (long long)(r0 + nr), (long long)s->R);
    GSB_REQUIRE(n0 >= 0 && nn >= 0 && n0 + nn <= s->N, "nodes [%lld, %lld) outside 0..%lld",
                (long long)n0, (long long)(n0 + nn), (long long)s->N);
    return GSB_OK;
}

int gsb_stack_upload_f32_async(gsb_stack* s, int64_t r0, int64_t nr, int64_t n0, int64_t nn,
                               const float* host, int64_t host_pitch, void* stream) {
    int rc = check_range(s, r0, nr, n0, nn);
    if (rc) return rc;
    GSB_REQUIRE(host && host_pitch >= nn, "bad host buffer");
    if (nr == 0 || nn == 0) return GSB_OK;
    GSB_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    GSB_CUDA(cudaMemcpy2DAsync(s->data + r0 * s->pitch + n0, (size_t)s->pitch * 4, host,
                               (size_t)host_pitch * 4, (size_t)nn * 4, (size_t)nr,
                               cudaMemcpyHostToDevice, st));
    return GSB_OK;
}

int gsb_stack_upload_f32(gsb_stack* s, int64_t r0, int64_t nr, int64_t n0, int64_t nn,
                         const float* host, int64_t host_pitch, void* stream) {
    int rc = gsb_stack_upload_f32_async(s, r0, nr, n0, nn, host, host_pitch, stream);
    if (rc) return rc;
    if (nr == 0 || nn == 0) return GSB_OK;
    GSB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return GSB_OK;
}

int gsb_stack_download_f32(const gsb_stack* s, int64_t r0, int64_t nr, int64_t n0, int64_t nn,
                           float* host, int64_t host_pitch, void* stream) {
    int rc = check_range(s, r0, nr, n0, nn);
    if (rc) return rc;
    GSB_REQUIRE(host && host_pitch >= nn, "bad host buffer");
    if (nr == 0 || nn == 0) return GSB_OK;
    GSB_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    GSB_CUDA(cudaMemcpy2DAsync(host, (size_t)host_pitch * 4, s->data + r0 * s->pitch + n0,
                               (size_t)s->pitch * 4, (size_t)nn * 4, (size_t)nr,
                               cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

int gsb_synth_fill(gsb_stack* s, uint64_t seed, int64_t dx, int64_t dy, int64_t dz, int64_t node0,
                   float nodata, void* stream) {
    GSB_REQUIRE(s, "stack is NULL");
    GSB_REQUIRE(dx > 0 && dy > 0 && dz > 0 && node0 >= 0 && node0 + s->N <= dx * dy * dz,
                "stack nodes [%lld, %lld) outside the %lldx%lldx%lld grid", (long long)node0,
                (long long)(node0 + s->N), (long long)dx, (long long)dy, (long long)dz);
    GSB_CUDA(cudaSetDevice(s->device));
    return gsb_launch_synth(s, seed, dx, dy, dz, node0, nodata, (cudaStream_t)stream);
}

// ------------------------------------------------------------------- K1
int gsb_decode_text(gsb_stack* s, int64_t r, const char* text, size_t nbytes, int header_lines,
                    int line_width, int64_t first_line, int64_t n0, int64_t nlines,
                    int64_t* bad_line, void* stream) {
    int rc = check_range(s, r, 1, n0, nlines);
    if (rc) return rc;
    GSB_REQUIRE(text || nbytes == 0, "text is NULL");
    GSB_REQUIRE(header_lines >= 0 && line_width >= 0 && first_line >= 0,
                "negative header_lines/line_width/first_line");
    if (bad_line) *bad_line = -1;
    if (nlines == 0) return GSB_OK;
    GSB_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    // skip the GSLIB header (name / nvars / varname lines, grid.py:250-258) on
    // the host: it is at most a few hundred bytes
    size_t body = 0;
    for (int h = 0; h < header_lines; ++h) {
        const void* nl = memchr(text + body, '\n', nbytes - body);
        if (!nl) { gsb_set_error("file has fewer than %d header lines", header_lines); return GSB_EPARSE; }
        body = (const char*)nl - text + 1;
    }
    const char* src = text + body;
    size_t len = nbytes - body;
    if (line_width > 0) {
        // only the byte range of the requested lines travels
        const size_t skip = (size_t)first_line * (size_t)line_width;
        const size_t need = (size_t)nlines * (size_t)line_width;
        GSB_REQUIRE(len >= skip + need,
                    "text too short: need %zu bytes after the header, have %zu", skip + need, len);
        src += skip;
        len = need;
        first_line = 0;
    }
    void* d = nullptr;
    rc = gsb_stage(s, len + 64, &d);
    if (rc) return rc;
    int64_t* d_bad = (int64_t*)d;
    char* dtext = (char*)d + 16;
    const int64_t minus1 = -1;
    GSB_CUDA(cudaMemcpyAsync(d_bad, &minus1, 8, cudaMemcpyHostToDevice, st));
    GSB_CUDA(cudaMemcpyAsync(dtext, src, len, cudaMemcpyHostToDevice, st));
    rc = gsb_launch_decode(s, r, dtext, len, 0, line_width, first_line, n0, nlines, d_bad, st);
    if (rc) return rc;
    int64_t hb = -1;
    GSB_CUDA(cudaMemcpyAsync(&hb, d_bad, 8, cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    if (hb >= 0) {
        if (bad_line) *bad_line = hb;
        gsb_set_error("could not convert string to float: data line %lld of realization %lld",
                      (long long)hb, (long long)r);
        return GSB_EPARSE;
    }
    return GSB_OK;
}

int gsb_decode_text_device(gsb_stack* s, int64_t r, const char* dtext, size_t nbytes,
                           int header_lines, int line_width, int64_t first_line, int64_t n0,
                           int64_t nlines, int64_t* d_bad_line, void* stream) {
    int rc = check_range(s, r, 1, n0, nlines);
    if (rc) return rc;
    GSB_REQUIRE(dtext, "text is NULL");
    GSB_CUDA(cudaSetDevice(s->device));
    void* d = nullptr;
    if (!d_bad_line) {        // a sink nobody reads
        rc = gsb_stage(s, 64, &d);
        if (rc) return rc;
        d_bad_line = (int64_t*)d;
    }
    return gsb_launch_decode(s, r, dtext, nbytes, header_lines, line_width, first_line, n0,
                             nlines, d_bad_line, (cudaStream_t)stream);
}

int gsb_encode_text_device(const gsb_stack* s, int64_t r, int64_t n0, int64_t nn, int width,
                           int prec, int crlf, char* dtext, void* stream) {
    int rc = check_range(s, r, 1, n0, nn);
    if (rc) return rc;
    GSB_REQUIRE(dtext, "text is NULL");
    GSB_CUDA(cudaSetDevice(s->device));
    return gsb_launch_encode(s, r, n0, nn, width, prec, crlf, dtext, (cudaStream_t)stream);
}

// ------------------------------------------------------------------- K2
int gsb_stats_grid_planes(uint32_t flags, int nq) {
    GsbPlanes pl;
    double dummy[GSB_MAX_PLANES] = {0};
    if (nq < 0 || nq > GSB_MAX_PLANES) return -1;
    if (gsb_build_planes(flags, 0.95, dummy, nq, &pl) != GSB_OK) return -1;
    return pl.n;
}

int gsb_stats_grid(gsb_stack* s, uint32_t flags, double p, const double* q, int nq, float nodata,
                   int64_t n0, int64_t nn, float* out, int64_t out_pitch, int out_on_device,
                   float* kernel_ms, void* stream) {
    int rc = check_range(s, 0, 0, n0, nn);
    if (rc) return rc;
    GSB_REQUIRE(nq >= 0 && (nq == 0 || q), "bad quantile list");
    GSB_REQUIRE(p > 0.0 && p <= 1.0, "p = %g outside (0, 1]", p);
    GsbPlanes pl;
    rc = gsb_build_planes(flags, p, q, nq, &pl);
    if (rc) return rc;
    if (pl.n == 0 || nn == 0) return GSB_OK;
    GSB_REQUIRE(out && out_pitch >= nn, "bad output buffer");
    GSB_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    float* d_out = out;
    int64_t d_pitch = out_pitch;
    if (!out_on_device) {
        void* d = nullptr;
        d_pitch = (nn + 3) & ~int64_t(3);
        rc = gsb_stage(s, (size_t)pl.n * (size_t)d_pitch * 4, &d);
        if (rc) return rc;
        d_out = (float*)d;
    }
    if (kernel_ms) GSB_CUDA(cudaEventRecord(s->ev0, st));
    rc = !pl.need_sort ? gsb_launch_k2_moments(s, pl, nodata, n0, nn, d_out, d_pitch, st)
       : s->R <= GSB_MAX_R_GRID ? gsb_launch_k2_sort(s, pl, nodata, n0, nn, d_out, d_pitch, st)
                                : gsb_launch_k2_deep(s, pl, nodata, n0, nn, d_out, d_pitch, st);
    if (rc) return rc;
    if (kernel_ms) GSB_CUDA(cudaEventRecord(s->ev1, st));
    if (!out_on_device) {
        GSB_CUDA(cudaMemcpy2DAsync(out, (size_t)out_pitch * 4, d_out, (size_t)d_pitch * 4,
                                   (size_t)nn * 4, (size_t)pl.n, cudaMemcpyDeviceToHost, st));
        GSB_CUDA(cudaStreamSynchronize(st));
    }
    if (kernel_ms) {
        GSB_CUDA(cudaEventSynchronize(s->ev1));
        GSB_CUDA(cudaEventElapsedTime(kernel_ms, s->ev0, s->ev1));
    }
    return GSB_OK;
}

int gsb_stats_columns(gsb_stack* s, const int64_t* node_ids, int S, int dz, int K, uint32_t flags,
                      double p, const double* q, int nq, float nodata, double* table,
                      int table_on_device, void* stream) {
    GSB_REQUIRE(s && node_ids && table, "NULL argument");
    GSB_REQUIRE(S >= 0 && dz >= 0 && K >= 1, "bad S/dz/K");
    GSB_REQUIRE(nq >= 0 && (nq == 0 || q), "bad quantile list");
    GsbPlanes pl;
    int rc = gsb_build_planes(flags, p, q, nq, &pl);
    if (rc) return rc;
    const int64_t ncols = (int64_t)S * dz;
    if (ncols == 0 || pl.n == 0) return GSB_OK;
    for (int64_t i = 0; i < ncols * K; ++i)
        GSB_REQUIRE(node_ids[i] >= -1 && node_ids[i] < s->N, "node id %lld outside the stack",
                    (long long)node_ids[i]);
    GSB_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t ids_bytes = ((size_t)ncols * K * 8 + 255) & ~size_t(255);
    const size_t tab_bytes = (size_t)ncols * pl.n * 8;
    void* d = nullptr;
    rc = gsb_stage(s, ids_bytes + tab_bytes, &d);
    if (rc) return rc;
    int64_t* d_ids = (int64_t*)d;
    double* d_tab = table_on_device ? table : (double*)((char*)d + ids_bytes);
    GSB_CUDA(cudaMemcpyAsync(d_ids, node_ids, (size_t)ncols * K * 8, cudaMemcpyHostToDevice, st));
    rc = gsb_launch_k2_columns(s, d_ids, S, dz, K, pl, nodata, d_tab, st);
    if (rc) return rc;
    if (!table_on_device) {
        GSB_CUDA(cudaMemcpyAsync(table, d_tab, tab_bytes, cudaMemcpyDeviceToHost, st));
        GSB_CUDA(cudaStreamSynchronize(st));
    }
    return GSB_OK;
}

int gsb_stats_columns_device(gsb_stack* s, const int64_t* d_node_ids, int S, int dz, int K,
                             uint32_t flags, double p, const double* q, int nq, float nodata,
                             double* d_table, void* stream) {
    GSB_REQUIRE(s && d_node_ids && d_table, "NULL argument");
    GSB_REQUIRE(S >= 0 && dz >= 0 && K >= 1, "bad S/dz/K");
    GSB_REQUIRE(nq >= 0 && (nq == 0 || q), "bad quantile list");
    GsbPlanes pl;
    int rc = gsb_build_planes(flags, p, q, nq, &pl);
    if (rc) return rc;
    if ((int64_t)S * dz == 0 || pl.n == 0) return GSB_OK;
    GSB_CUDA(cudaSetDevice(s->device));
    return gsb_launch_k2_columns(s, d_node_ids, S, dz, K, pl, nodata, d_table, (cudaStream_t)stream);
}

int gsb_detect_table(const double* d_clim, const double* d_table, const double* d_fixtab,
                     const int64_t* d_zoff, int64_t sstride, int C, int Cfix, int c_mean,
                     int c_lperc, int c_rperc, int c_fa, int c_fb, int c_fc, int S, int dz,
                     int method, double thr, double nodata, double* d_out, int32_t* d_detected,
                     int device, void* stream) {
    GSB_REQUIRE(d_clim && d_table && d_fixtab && d_zoff && d_out && d_detected, "NULL argument");
    GSB_REQUIRE(method >= GSB_METHOD_MEAN && method <= GSB_METHOD_PERCENTILE,
                "Method %d invalid or incomplete.", method);
    GSB_REQUIRE(S >= 0 && dz >= 0 && C >= 1 && Cfix >= 1, "bad S/dz/C");
    GSB_REQUIRE(c_mean >= 0 && c_mean < C && c_lperc >= 0 && c_lperc < C && c_rperc >= 0 &&
                c_rperc < C && c_fa >= 0 && c_fa < Cfix && c_fb >= 0 && c_fb < Cfix && c_fc < C,
                "column index outside the table");
    GSB_REQUIRE(method != GSB_METHOD_SKEWNESS || c_fc >= 0, "skewness method needs the skewness column");
    if (S == 0 || dz == 0) return GSB_OK;
    GSB_CUDA(cudaSetDevice(device));
    return gsb_launch_k3_detect_table(d_clim, d_table, d_fixtab, d_zoff, sstride, C, Cfix, c_mean,
                                      c_lperc, c_rperc, c_fa, c_fb, c_fc, S, dz, method, thr,
                                      nodata, d_out, d_detected, (cudaStream_t)stream);
}

int gsb_extract_columns(gsb_stack* s, const int64_t* node_ids, int S, int dz, int K,
                        float* values, void* stream) {
    GSB_REQUIRE(s && node_ids && values, "NULL argument");
    GSB_REQUIRE(S >= 0 && dz >= 0 && K >= 1, "bad S/dz/K");
    const int64_t ncols = (int64_t)S * dz;
    if (ncols == 0) return GSB_OK;
    for (int64_t i = 0; i < ncols * K; ++i)
        GSB_REQUIRE(node_ids[i] >= -1 && node_ids[i] < s->N, "node id %lld outside the stack",
                    (long long)node_ids[i]);
    GSB_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t ids_bytes = ((size_t)ncols * K * 8 + 255) & ~size_t(255);
    const size_t val_bytes = (size_t)ncols * K * (size_t)s->R * 4;
    void* d = nullptr;
    int rc = gsb_stage(s, ids_bytes + val_bytes, &d);
    if (rc) return rc;
    int64_t* d_ids = (int64_t*)d;
    float* d_val = (float*)((char*)d + ids_bytes);
    GSB_CUDA(cudaMemcpyAsync(d_ids, node_ids, (size_t)ncols * K * 8, cudaMemcpyHostToDevice, st));
    rc = gsb_launch_extract(s, d_ids, ncols, K, d_val, st);
    if (rc) return rc;
    GSB_CUDA(cudaMemcpyAsync(values, d_val, val_bytes, cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

// ------------------------------------------------------------------- K3
int gsb_detect(const double* clim, const double* mean, const double* lperc, const double* rperc,
               const double* fix_a, const double* fix_b, const double* fix_c, int S, int dz,
               int method, double thr, double nodata, double* filled, double* homog, double* flag,
               int32_t* detected, int on_device, int device, void* stream) {
    GSB_REQUIRE(clim && mean && lperc && rperc && fix_a && filled && homog && flag && detected,
                "NULL argument");
    GSB_REQUIRE(method >= GSB_METHOD_MEAN && method <= GSB_METHOD_PERCENTILE,
                "Method %d invalid or incomplete.", method);
    GSB_REQUIRE(method != GSB_METHOD_SKEWNESS || (fix_b && fix_c), "skewness method needs mean and skewness columns");
    GSB_REQUIRE(method != GSB_METHOD_PERCENTILE || fix_b, "percentile method needs both bounds");
    GSB_REQUIRE(S >= 0 && dz >= 0, "bad S/dz");
    if (S == 0 || dz == 0) return GSB_OK;
    GSB_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    if (on_device)
        return gsb_launch_k3_detect(clim, mean, lperc, rperc, fix_a, fix_b, fix_c, S, dz, method,
                                    thr, nodata, filled, homog, flag, detected, st);
    // host buffers: 7 inputs + 3 outputs + counts in one staging block that is
    // kept per device (grow-only).  A stream-ordered cudaMallocAsync/cudaFreeAsync
    // pair here returned its memory to the OS at the synchronize below, and that
    // release waited for kernels running on OTHER streams (the persistent K2
    // kernel): 7 ms per call instead of 0.1 ms.
    const size_t n = (size_t)S * dz, nb = n * 8;
    const size_t need = nb * 10 + (size_t)S * 4 + 64;
    static std::mutex mu;
    static void* cache_dev[64] = {nullptr};
    static void* cache_pin[64] = {nullptr};
    static size_t cache_bytes[64] = {0};
    GSB_REQUIRE(device >= 0 && device < 64, "bad device %d", device);
    std::lock_guard<std::mutex> lock(mu);
    if (cache_bytes[device] < need) {
        if (cache_dev[device]) cudaFree(cache_dev[device]);
        if (cache_pin[device]) cudaFreeHost(cache_pin[device]);
        cache_dev[device] = cache_pin[device] = nullptr;
        cache_bytes[device] = 0;
        const size_t cap = (need + 65535) & ~size_t(65535);
        GSB_CUDA(cudaMalloc(&cache_dev[device], cap));
        GSB_CUDA(cudaMallocHost(&cache_pin[device], cap));
        cache_bytes[device] = cap;
    }
    char* d = (char*)cache_dev[device];
    char* h = (char*)cache_pin[device];
    const double* hin[7] = {clim, mean, lperc, rperc, fix_a, fix_b ? fix_b : fix_a,
                            fix_c ? fix_c : fix_a};
    double* din[7];
    for (int i = 0; i < 7; ++i) {
        din[i] = (double*)(d + nb * i);
        memcpy(h + nb * i, hin[i], nb);
    }
    GSB_CUDA(cudaMemcpyAsync(d, h, nb * 7, cudaMemcpyHostToDevice, st));   // one pinned H2D
    double* dfilled = (double*)(d + nb * 7);
    int32_t* ddet = (int32_t*)(d + nb * 10);
    int rc = gsb_launch_k3_detect(din[0], din[1], din[2], din[3], din[4], din[5], din[6], S, dz,
                                  method, thr, nodata, dfilled, (double*)(d + nb * 8),
                                  (double*)(d + nb * 9), ddet, st);
    if (rc != GSB_OK) return rc;
    // outputs are contiguous behind the inputs: one pinned D2H
    GSB_CUDA(cudaMemcpyAsync(h + nb * 7, d + nb * 7, nb * 3 + (size_t)S * 4, cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    memcpy(filled, h + nb * 7, nb);
    memcpy(homog, h + nb * 8, nb);
    memcpy(flag, h + nb * 9, nb);
    memcpy(detected, h + nb * 10, (size_t)S * 4);
    return rc;
}

}  // extern "C"
