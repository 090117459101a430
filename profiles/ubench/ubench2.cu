// ubench2.cu -- round-2 micro-benchmarks that size the tile-resident K2 design:
//   (1) how fast one SM can pull [R rows x 16|32 nodes] column tiles, by TMA or
//       by plain coalesced LDG + STS, with 1 / 2 / 3 CTAs resident per SM and
//       0 or 3 read passes over the tile (the K2 passes);
//   (2) shared-memory RED / ATOMS / STS cost when the 32 lanes of a warp own 16
//       columns x 2 rows and the per-column cells are interleaved
//       (word = bin * 16 + column): at most 2-way conflicts by construction;
//   (3) FP64 add/fma issue rate.
// Standalone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
//        -o profiles/ubench/ubench2 profiles/ubench/ubench2.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
    printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

constexpr int THREADS = 512;

// ------------------------------------------------------------------ tile streaming
// MODE 0: TMA boxes {COLS, BOXR}; MODE 1: LDG (coalesced rows) + STS
template <int COLS, int MODE>
__global__ void __launch_bounds__(THREADS, 3)
k_tiles(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ g, long long pitch,
        int R, int boxr, long long ntiles, int passes, float* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* tile = reinterpret_cast<float*>(smem);
    const int rows_pad = (R + 63) & ~63;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)rows_pad * COLS * 4);
    const int nbox = (R + boxr - 1) / boxr;
    auto issue = [&](long long t) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"((uint32_t)(nbox * boxr * COLS * 4)) : "memory");
        for (int b = 0; b < nbox; ++b)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(s32(smem + (size_t)b * boxr * COLS * 4)), "l"(&tmap), "r"(s32(bar)), "r"((int)(t * COLS)), "r"(b * boxr) : "memory");
    };
    if (MODE == 0 && threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (blockIdx.x < ntiles) issue(blockIdx.x);
    }
    __syncthreads();
    uint32_t phase = 0;
    float acc = 0.f;
    const int tid = threadIdx.x;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        if (MODE == 0) {
            asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(bar)), "r"(phase) : "memory");
            phase ^= 1;
        } else {
            // each warp-instruction reads 128 contiguous bytes of one (COLS=32) or two
            // (COLS=16) rows; 16 loads in flight per thread
            const float* gp = g + t * COLS;
            constexpr int RPI = THREADS / COLS;           // rows per CTA-wide load instruction
            const int c = tid % COLS, r0 = tid / COLS;
            for (int rb = 0; rb < R; rb += RPI * 16) {
                float v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int r = rb + u * RPI + r0;
                    v[u] = r < R ? __ldg(gp + (long long)r * pitch + c) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int r = rb + u * RPI + r0;
                    if (r < rows_pad) tile[r * COLS + c] = v[u];
                }
            }
            __syncthreads();
        }
        // passes: lane = (column pair, row in a group of 64/COLS*... rows), LDS.64
        for (int p = 0; p < passes; ++p) {
            constexpr int CP = COLS / 2;
            const int cp = tid % CP, rr = tid / CP;
            constexpr int RPI = THREADS / CP;
            for (int r = rr; r < R; r += RPI) {
                const float2 v = *reinterpret_cast<const float2*>(tile + r * COLS + cp * 2);
                acc += v.x * v.y + (float)p;
            }
        }
        if (passes == 0) acc += tile[(tid * 33) % (R * COLS)];
        __syncthreads();
        if (MODE == 0 && threadIdx.x == 0) { const long long nx = t + gridDim.x; if (nx < ntiles) issue(nx); }
    }
    if (acc == 1234.5f) out[0] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int COLS, int MODE>
void run_tiles(EncodeFn enc, float* d, long long N, int R, int sms, int occ, int boxr, int passes,
               CUtensorMapL2promotion prom, const char* promname, float* d_out) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)R};
    cuuint64_t strides[1] = {(cuuint64_t)N * 4};
    cuuint32_t box[2] = {(cuuint32_t)COLS, (cuuint32_t)boxr};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return; }
    const int rows_pad = (R + 63) & ~63;
    size_t smem = (size_t)rows_pad * COLS * 4 + 64;
    // pad the request so that exactly `occ` CTAs fit per SM
    const size_t per = (227 * 1024) / occ - 1024;
    if (smem > per) { printf("tile does not fit at occ %d\n", occ); return; }
    if (occ < 3) smem = per > smem ? per : smem;
    CK(cudaFuncSetAttribute(k_tiles<COLS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long ntiles = N / COLS;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_tiles<COLS, MODE><<<sms * occ, THREADS, smem>>>(tm, d, N, R, boxr, ntiles, passes, d_out);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    printf("%s tile [%d x %2d] occ %d boxr %3d passes %d prom %-4s: %7.3f ms %7.1f GB/s\n", MODE ? "LDG" : "TMA", R, COLS,
           occ, boxr, passes, promname, best, (double)N * R * 4 / best / 1e6);
    fflush(stdout);
}

// ------------------------------------------------------------------ smem op cost
enum { O_RED_IL, O_ATOM_IL, O_STS_IL, O_RED_RND, O_LDS_RND_DEP, O_LDS64_LIN_DEP, O_DADD, O_DFMA, O_RED16_IL, O_COUNT };
const char* onames[] = {"RED.ADD interleaved (16 col x 2 rows/warp)", "ATOMS.ADD ret interleaved", "STS interleaved",
    "RED.ADD random (one column per warp)", "LDS random, consumed", "LDS.64 linear, consumed", "DADD (4 chains)", "DFMA (4 chains)",
    "RED.ADD 16-bit cells interleaved"};

template <int T>
__global__ void __launch_bounds__(THREADS, 1) k_ops2(int iters, long long* cyc, uint32_t* sink) {
    extern __shared__ uint32_t sm[];          // 512 bins x 16 columns words = 32 KB
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 8192; i += THREADS) sm[i] = 0;
    uint32_t idx[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const uint32_t r = hash32((blockIdx.x * THREADS + threadIdx.x) * 32 + j + 777u);
        const uint32_t bin = r & 511;
        if (T == O_RED_RND || T == O_LDS_RND_DEP) idx[j] = ((threadIdx.x >> 5) & 7) * 1024 + (r & 1023);
        else if (T == O_LDS64_LIN_DEP) idx[j] = ((j * 32 + lane) * 2) & 8191;
        else if (T == O_RED16_IL) idx[j] = (r & 1023);          // 1024 16-bit bins
        else idx[j] = bin * 16 + (lane & 15);
    }
    double dacc[4] = {1.0, 2.0, 3.0, 4.0};
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if constexpr (T == O_RED_IL || T == O_RED_RND) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(s32(sm + idx[j])) : "memory");
        } else if constexpr (T == O_RED16_IL) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const uint32_t b = idx[j];
                const uint32_t w = (b >> 1) * 16 + (lane & 15);
                asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(s32(sm + w)), "r"(1u << ((b & 1) << 4)) : "memory");
            }
        } else if constexpr (T == O_ATOM_IL) {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc += atomicAdd(sm + idx[j], 1u);
        } else if constexpr (T == O_STS_IL) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(s32(sm + idx[j])), "r"(acc + j) : "memory");
        } else if constexpr (T == O_LDS_RND_DEP) {
            uint32_t v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j)
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[j]) : "r"(s32(sm + idx[j]) + (acc & 4u)) : "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) acc += v[j];
        } else if constexpr (T == O_LDS64_LIN_DEP) {
            uint2 v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j)
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v[j].x), "=r"(v[j].y) : "r"(s32(sm + idx[j]) + (acc & 8u)) : "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) acc += v[j].x ^ v[j].y;
        } else if constexpr (T == O_DADD) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int j = 0; j < 4; ++j) dacc[j] = __dadd_rn(dacc[j], 1.000001);
        } else if constexpr (T == O_DFMA) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int j = 0; j < 4; ++j) dacc[j] = __fma_rn(dacc[j], 1.000001, 0.5);
        }
    }
    const long long t1 = clock64();
#pragma unroll
    for (int j = 0; j < 4; ++j) acc += (uint32_t)__double2loint(dacc[j]);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc + sm[lane];
}

template <int T>
void run_op2(int sms, int iters, long long* d_cyc, uint32_t* d_sink) {
    const size_t smem = 8192 * 4;
    k_ops2<T><<<sms, THREADS, smem>>>(10, d_cyc, d_sink);
    k_ops2<T><<<sms, THREADS, smem>>>(iters, d_cyc, d_sink);
    CK(cudaDeviceSynchronize());
    long long* h = (long long*)malloc(sms * sizeof(long long));
    CK(cudaMemcpy(h, d_cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += h[i];
    avg /= sms;
    const double per = avg / ((double)iters * 32 * (THREADS / 32));
    printf("%-46s %8.3f cyc/warp-instr/SM  %7.2f lanes/clk/SM\n", onames[T], per, 32.0 / per);
    free(h);
}

int main(int argc, char** argv) {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    long long* d_cyc; uint32_t* d_sink;
    CK(cudaMalloc(&d_cyc, sms * sizeof(long long)));
    CK(cudaMalloc(&d_sink, 1024));
    const int it = 1000;
    run_op2<O_RED_IL>(sms, it, d_cyc, d_sink);
    run_op2<O_RED16_IL>(sms, it, d_cyc, d_sink);
    run_op2<O_ATOM_IL>(sms, it, d_cyc, d_sink);
    run_op2<O_STS_IL>(sms, it, d_cyc, d_sink);
    run_op2<O_RED_RND>(sms, it, d_cyc, d_sink);
    run_op2<O_LDS_RND_DEP>(sms, it, d_cyc, d_sink);
    run_op2<O_LDS64_LIN_DEP>(sms, it, d_cyc, d_sink);
    run_op2<O_DADD>(sms, it, d_cyc, d_sink);
    run_op2<O_DFMA>(sms, it, d_cyc, d_sink);

    EncodeFn enc = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr));
    const long long N = 2000000; const int R = 1000;       // 8 GB
    float* d; CK(cudaMalloc(&d, (size_t)N * R * 4));
    CK(cudaMemset(d, 0, (size_t)N * R * 4));
    float* out = (float*)d_sink;
    const CUtensorMapL2promotion P256 = CU_TENSOR_MAP_L2_PROMOTION_L2_256B, P128 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 P0 = CU_TENSOR_MAP_L2_PROMOTION_NONE;
    run_tiles<32, 0>(enc, d, N, R, sms, 1, 256, 0, P256, "256", out);
    run_tiles<32, 0>(enc, d, N, R, sms, 1, 256, 3, P256, "256", out);
    run_tiles<16, 0>(enc, d, N, R, sms, 1, 256, 0, P256, "256", out);
    run_tiles<16, 0>(enc, d, N, R, sms, 2, 256, 0, P256, "256", out);
    run_tiles<16, 0>(enc, d, N, R, sms, 3, 256, 0, P256, "256", out);
    run_tiles<16, 0>(enc, d, N, R, sms, 2, 256, 3, P256, "256", out);
    run_tiles<16, 0>(enc, d, N, R, sms, 3, 256, 3, P256, "256", out);
    run_tiles<16, 0>(enc, d, N, R, sms, 2, 64, 0, P256, "256", out);
    run_tiles<16, 0>(enc, d, N, R, sms, 2, 256, 0, P128, "128", out);
    run_tiles<16, 0>(enc, d, N, R, sms, 2, 256, 0, P0, "none", out);
    run_tiles<16, 0>(enc, d, N, R, sms, 3, 256, 0, P0, "none", out);
    run_tiles<16, 1>(enc, d, N, R, sms, 1, 256, 0, P256, "-", out);
    run_tiles<16, 1>(enc, d, N, R, sms, 2, 256, 0, P256, "-", out);
    run_tiles<16, 1>(enc, d, N, R, sms, 3, 256, 0, P256, "-", out);
    run_tiles<16, 1>(enc, d, N, R, sms, 2, 256, 3, P256, "-", out);
    run_tiles<16, 1>(enc, d, N, R, sms, 3, 256, 3, P256, "-", out);
    run_tiles<32, 1>(enc, d, N, R, sms, 1, 256, 0, P256, "-", out);
    return 0;
}
