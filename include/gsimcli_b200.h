/* gsimcli_b200.h -- C ABI of libgsimcli_b200.so
 *
 * B200-native (sm_100a) implementation of ONE hot path of iled/gsimcli: the
 * reduction of a stack of DSS realizations to per-node local PDFs, and the
 * per-station detect/correct step.  This header is the drop-in boundary: the
 * entry points are what a Python (ctypes) binding for the reference's
 * `tools/grid.py::GridFiles` and `tools/homog.py::detect` would call.  All
 * reference citations are relative to the reference checkout
 * (`GSIMCLI/tools/...`).  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *   - every function returns an int status (GSB_OK == 0); a message for the
 *     last failing call on the calling thread is available from
 *     gsb_last_error().  There is NO CPU fallback: without a CUDA device every
 *     compute entry point returns GSB_ECUDA.
 *   - the library owns device memory inside the opaque `gsb_stack`; callers own
 *     every host buffer; nothing library-allocated is ever returned.
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = default stream).
 *     Calls taking HOST output pointers synchronise `stream` before returning;
 *     calls taking DEVICE pointers only enqueue work.
 *   - the stack is `[R][pitch]` float32, realization-major: one realization
 *     (one GSLIB file, x fastest, then y, then z -- grid.py:254-258) is one
 *     contiguous row, so a warp owning 32 consecutive nodes reads 128
 *     contiguous bytes of every row.  A stack may hold only the node range
 *     `[node0, node0+N)` of the grid (multi-GPU node sharding).
 */
#ifndef GSIMCLI_B200_H
#define GSIMCLI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gsb_stack gsb_stack;

/* status codes; the Python shim maps them to the reference's exceptions:
 * ENOENT -> IOError('File ... not found.') (grid.py:506,574); EINVAL ->
 * ValueError (homog.py:153); EPARSE -> ValueError, like float() on a malformed
 * line (grid.py:673,823); ECUDA/ENOMEM -> RuntimeError/MemoryError. */
enum {
    GSB_OK = 0,
    GSB_ENOENT = 1,
    GSB_EINVAL = 2,
    GSB_EPARSE = 3,
    GSB_ECUDA = 4,
    GSB_ENCCL = 5,   /* reserved: the library exports no collective.  The only exchange of
                        the path -- the all-gather of the per-station tables of a
                        node-sharded stack -- is the binding's job (torch.distributed
                        on the device buffers gsb_stats_columns_device fills) */
    GSB_ENOMEM = 6
};

/* statistic selectors.  Bit order == column order of the stats_area table
 * (grid.py:884-912): mean, median, skewness, variance, std, coefvar,
 * lperc+rperc.  GSB_MODE is an extension (not in the reference). */
enum {
    GSB_MEAN = 1,     /* lmean     bn.nanmean            grid.py:683,832 */
    GSB_MEDIAN = 2,   /* lmed      bn.nanmedian          grid.py:685,834 */
    GSB_SKEW = 4,     /* lskew     pd.Series.skew (G1)   grid.py:687,836 */
    GSB_VAR = 8,      /* lvar      bn.nanvar(ddof=1)     grid.py:689,838 */
    GSB_STD = 16,     /* lstd      bn.nanstd(ddof=1)     grid.py:691,840 */
    GSB_COEFVAR = 32, /* lcoefvar  std/mean*100          grid.py:692-698 */
    GSB_PERC = 64,    /* lperc     quantile pair         grid.py:699-701 */
    GSB_MODE = 128    /* extension: most frequent value, ties -> smallest */
};

/* correction methods of detect() (homog.py:136-153, 215-233) */
enum {
    GSB_METHOD_MEAN = 0,
    GSB_METHOD_MEDIAN = 1,
    GSB_METHOD_SKEWNESS = 2,
    GSB_METHOD_PERCENTILE = 3
};

#define GSB_MAX_PLANES 64
#define GSB_MAX_R_GRID 1024 /* deepest column of the tile kernels; deeper stacks take the
                               column kernel node by node (slower, same results) */

int gsb_version(void);
/* copies the calling thread's last error message into buf (NUL terminated) */
int gsb_last_error(char* buf, size_t n);
int gsb_device_count(int* count);
/* sm_count, total memory and the device name (name may be NULL) */
int gsb_device_info(int device, int* sm_count, int* cc_major, int* cc_minor,
                    size_t* total_mem, char* name, size_t name_len);

/* ---- stack lifecycle ---------------------------------------------------
 * GridFiles.load / open_files (grid.py:449-574) create the stack that the
 * reference keeps as R open file handles; GridFiles.dump (grid.py:583-589)
 * destroys it.  R = number of realizations (nfiles), N = nodes held. */
int gsb_stack_create(int device, int64_t R, int64_t N, gsb_stack** out);
int gsb_stack_destroy(gsb_stack* s);
int gsb_stack_info(const gsb_stack* s, int64_t* R, int64_t* N, int64_t* pitch,
                   int* device);
/* tuning knobs.  "k2_sm_reserve" = number of SMs the full-grid kernel leaves
 * free (default 0) so that station-column / detect launches on another stream
 * run concurrently with it instead of queueing behind its persistent CTAs.
 * Tests and profiling only: "k2_algo" selects the full-grid kernel (0 = default,
 * 1 = counting sort, 2 = register bitonic for every R, 3 = register tile, 4 = the
 * experimental staged-select kernel for the reference's call shape -- other shapes
 * take the register tile), "k2_nw" the warps per CTA of the register-tile /
 * staged-select kernel, "k2_stats" enables gsb_debug_cs_counters, "k1_algo" = 1
 * sends '%12.4f CRLF' text to the experimental 8-records-per-thread decoder. */
int gsb_stack_set_option(gsb_stack* s, const char* key, int64_t value);
/* diagnostic: number of 32-node tiles of the last gsb_stats_grid call on this
 * stack that the register-tile kernel handed to the counting-sort kernel
 * (synchronises the device). */
int gsb_debug_k2_flagged(gsb_stack* s, int* count);
/* raw device pointer of row 0 (float32, row pitch `pitch` elements): the
 * tensor handle torch wraps for NCCL plumbing.  Never dereference on host. */
int gsb_stack_device_ptr(const gsb_stack* s, void** dptr);

/* ---- filling the stack -------------------------------------------------
 * upload/download rows [r0, r0+nr) x nodes [n0, n0+nn) from/to a host float32
 * matrix with row pitch `host_pitch` elements. */
int gsb_stack_upload_f32(gsb_stack* s, int64_t r0, int64_t nr, int64_t n0,
                         int64_t nn, const float* host, int64_t host_pitch,
                         void* stream);
/* enqueue-only form of the upload: `host` must be pinned and stay untouched
 * until `stream` has passed the copy (the synchronous form above waits). */
int gsb_stack_upload_f32_async(gsb_stack* s, int64_t r0, int64_t nr,
                               int64_t n0, int64_t nn, const float* host,
                               int64_t host_pitch, void* stream);
int gsb_stack_download_f32(const gsb_stack* s, int64_t r0, int64_t nr,
                           int64_t n0, int64_t nn, float* host,
                           int64_t host_pitch, void* stream);

/* K1 -- GSLIB ASCII grid decoder.  Replaces `arr[i] = grid.readline()`
 * (grid.py:673), `float(a)` (grid.py:822-823), skip_lines (utils.py:66-78)
 * and np.loadtxt(skiprows) (grid.py:326).  `text` is the HOST image of one
 * realization file; data lines [first_line, first_line+nlines) (0-based,
 * counted after the `header_lines` header lines) are decoded into row r,
 * nodes [n0, n0+nlines) -- a node-sharded stack passes first_line = its
 * first grid node.  `line_width` > 0 asserts fixed-width records
 * (every line exactly that many bytes incl. EOL; 14 for '%12.4f\r\n',
 * interface/gui.py:1414) and enables O(1) line addressing; 0 = general form
 * (newline index built on the device).  Python float() syntax: blanks, CR,
 * sign, digits, '.', e/E exponent, inf/nan.  Conversion is correctly rounded
 * (decimal -> float64 -> float32, the reference's float() then our storage
 * cast).  A malformed line gives GSB_EPARSE and *bad_line = its 0-based
 * index (like ValueError from float(), grid.py:823). */
int gsb_decode_text(gsb_stack* s, int64_t r, const char* text, size_t nbytes,
                    int header_lines, int line_width, int64_t first_line,
                    int64_t n0, int64_t nlines, int64_t* bad_line,
                    void* stream);
/* same, `text` already resident in device memory (enqueue only; *bad_line is
 * a DEVICE int64 initialised to -1 by the caller, may be NULL) */
int gsb_decode_text_device(gsb_stack* s, int64_t r, const char* dtext,
                           size_t nbytes, int header_lines, int line_width,
                           int64_t first_line, int64_t n0, int64_t nlines,
                           int64_t* d_bad_line, void* stream);
/* inverse of K1 for synthetic inputs and GridArr.save-style output
 * (grid.py:341-347): row r, nodes [n0, n0+nn) -> fixed-width records
 * ('%<width>.<prec>f' + eol, correctly rounded like printf; "nan" / "inf" for
 * non-finite values) written to a DEVICE buffer of nn*(width+eol_len) bytes.
 * `crlf` bit 0: EOL is "\r\n" instead of "\n"; bit 1: left-justified
 * ('%-<width>.<prec>f', the np.savetxt format of GridArr.save). */
int gsb_encode_text_device(const gsb_stack* s, int64_t r, int64_t n0,
                           int64_t nn, int width, int prec, int crlf,
                           char* dtext, void* stream);

/* seeded synthetic fill (bit-identical to gsimcli_b200/synth.py): the stack
 * holds nodes [node0, node0+N) of a dx*dy*dz grid. */
int gsb_synth_fill(gsb_stack* s, uint64_t seed, int64_t dx, int64_t dy,
                   int64_t dz, int64_t node0, float nodata, void* stream);

/* ---- K2: realization-axis reduction, full grid --------------------------
 * GridFiles.stats (grid.py:601-735) for nodes [n0, n0+nn) of the stack.
 * Output planes, each `nn` float32, written to `out + k*out_pitch` in this
 * order: the maps selected by `flags` in bit order (GSB_PERC contributes two
 * planes, quantiles (1-p)/2 and 1-(1-p)/2, grid.py:700-701), then the `nq`
 * extra quantiles `q[]` (extension: percentile lists), then GSB_MODE last.
 * `nodata` is compared in float32, the precision the stack stores
 * (bn.replace, grid.py:681).  out_on_device: 0 = host pointer (D2H + sync
 * inside the call), 1 = device pointer (enqueue only).
 * *kernel_ms (may be NULL, host output only) receives the CUDA-event time of
 * the reduction kernel on `stream`. */
int gsb_stats_grid(gsb_stack* s, uint32_t flags, double p, const double* q,
                   int nq, float nodata, int64_t n0, int64_t nn, float* out,
                   int64_t out_pitch, int out_on_device, float* kernel_ms,
                   void* stream);
/* number of planes gsb_stats_grid writes for (flags, nq) */
int gsb_stats_grid_planes(uint32_t flags, int nq);

/* ---- K2c: station columns ----------------------------------------------
 * GridFiles.stats_area (grid.py:737-918) for S stations: per station and
 * time layer, statistics over the R*K values of the disc neighbourhood.
 * node_ids is HOST int64 [S][dz][K] of node indices LOCAL to the stack
 * (line number - 1 - node0), -1 = absent (outside this stack / the grid).
 * Values are gathered in the reference's order arr[i + j*K] (grid.py:823),
 * promoted to float64, and reduced in float64 (sequential sums, like
 * bottleneck).  table is HOST or DEVICE float64 [S][dz][C], C columns in
 * stats_area order (mean, median, skewness, variance, std, coefvar, lperc,
 * rperc -- those selected), then `nq` extra quantiles, then mode. */
int gsb_stats_columns(gsb_stack* s, const int64_t* node_ids, int S, int dz,
                      int K, uint32_t flags, double p, const double* q, int nq,
                      float nodata, double* table, int table_on_device,
                      void* stream);
/* same reduction, everything resident: d_node_ids and d_table are DEVICE
 * pointers, nothing is validated or copied, the call only enqueues.  This is
 * the step form of the station path: the ids of a network's stations are
 * uploaded once, the table stays on the device for gsb_detect_table (and for
 * the NCCL gather of a node-sharded stack). */
int gsb_stats_columns_device(gsb_stack* s, const int64_t* d_node_ids, int S,
                             int dz, int K, uint32_t flags, double p,
                             const double* q, int nq, float nodata,
                             double* d_table, void* stream);
/* raw extraction at station coordinates (`save=True`, grid.py:851-867):
 * values HOST float32 [S][dz][K][R] in realization order. */
int gsb_extract_columns(gsb_stack* s, const int64_t* node_ids, int S, int dz,
                        int K, float* values, void* stream);

/* ---- K3: fused per-station detection / correction ------------------------
 * homog.py:191-239 for S stations x dz time steps.  All arrays HOST or all
 * DEVICE (on_device).  Inputs: clim [S][dz] observed values with NaN =
 * missing (already nodata-replaced, homog.py:188); mean, lperc, rperc [S][dz]
 * local-PDF columns; fix_a, fix_b, fix_c [S][dz] meaning by method:
 *   MEAN       fix_a = mean
 *   MEDIAN     fix_a = median
 *   SKEWNESS   fix_a = median, fix_b = mean, fix_c = skewness; threshold thr
 *              (np.where(skew > thr, median, mean), homog.py:219-221)
 *   PERCENTILE fix_a = rperc', fix_b = lperc' of the correction percentile
 *              (np.where(clim > rperc, rperc', lperc'), homog.py:232-233)
 * Outputs: filled [S][dz] (NaN -> mean, homog.py:200-201), homog [S][dz]
 * (clim.where(~hom_where, fix), :235), flag [S][dz] (observed value where
 * detected else `nodata`, :237-238), detected [S] (hom_where.sum(), :206).
 * Detection is ~between(lperc, rperc), inclusive both ends, NaN -> detected. */
int gsb_detect(const double* clim, const double* mean, const double* lperc,
               const double* rperc, const double* fix_a, const double* fix_b,
               const double* fix_c, int S, int dz, int method, double thr,
               double nodata, double* filled, double* homog, double* flag,
               int32_t* detected, int on_device, int device, void* stream);

/* K3 reading the station tables where gsb_stats_columns_device (and, on a
 * node-sharded stack, the all-gather of the per-rank slabs) left them; every
 * pointer is DEVICE, the call only enqueues.  Row (s, z) of `d_table` (C
 * columns) and `d_fixtab` (Cfix columns; the same table unless the percentile
 * method corrects with another interval, homog.py:222-233) starts at element
 * (d_zoff[z] + s * sstride) * C: a plain [S][dz][C] table has d_zoff[z] = z,
 * sstride = dz; gathered slabs [world][S][width][C] have d_zoff[z] =
 * rank(z) * S * width + z - z0(rank(z)), sstride = width.  c_* are column
 * indices: c_mean, c_lperc, c_rperc in d_table; c_fa, c_fb in d_fixtab with
 * the meaning of fix_a / fix_b above; c_fc (skewness, -1 when unused) in
 * d_table.  d_out is [3][S][dz]: filled, homog, flag. */
int gsb_detect_table(const double* d_clim, const double* d_table,
                     const double* d_fixtab, const int64_t* d_zoff,
                     int64_t sstride, int C, int Cfix, int c_mean, int c_lperc,
                     int c_rperc, int c_fa, int c_fb, int c_fc, int S, int dz,
                     int method, double thr, double nodata, double* d_out,
                     int32_t* d_detected, int device, void* stream);

/* ---- diagnostics ----------------------------------------------------------
 * Counters of the counting-sort K2 kernel, accumulated only while the stack
 * option "k2_stats" is non-zero: out[0] columns reduced,
 * [1] columns that held an over-full bin, [2] columns that took the generic
 * shared-memory sort, [3] sum of odd-even phases, [4] over-full bins sorted, [5] SM cycles warps
 * spent waiting for a tile, [6] SM cycles of all warps.
 * `reset` != 0 clears them.  Not part of the reference interface. */
int gsb_debug_cs_counters(unsigned long long* out, int reset);

#ifdef __cplusplus
}
#endif
#endif /* GSIMCLI_B200_H */
