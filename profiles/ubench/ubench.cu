// ubench.cu -- SM-level micro-benchmarks that size the K2 design choices
// (shared-memory atomics / random LDS+STS / shuffles / FMNMX / MATCH, and the
// TMA streaming rate of [R][16|32 nodes] column tiles).  Standalone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
//        -o profiles/ubench/ubench profiles/ubench/ubench.cu -lcuda
// Output: one line per test, cycles per warp-instruction per SM (all 148 SMs
// busy, 16 warps per SM) and the implied lanes/clk/SM.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
    printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int WARPS = 16;
constexpr int THREADS = WARPS * 32;
constexpr int NELEM = 32;      // per-lane independent items (like one column)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

enum { T_ATOM_NORET, T_ATOM_RET, T_ATOM_OR, T_STS, T_LDS, T_LDS_LIN, T_ATOM_LIN,
       T_SHFL, T_FMNMX, T_FFMA, T_MATCH, T_F2I, T_ATOM_RET_2K, T_STS_LDS_MIX, T_VIMNMX, T_COUNT };
const char* names[] = {"ATOMS.ADD noret random(1024w/warp)", "ATOMS.ADD ret random(1024w/warp)",
    "ATOMS.OR ret random(1024w/warp)", "STS random(1024w/warp)", "LDS random(1024w/warp)",
    "LDS linear (conflict-free)", "ATOMS.ADD ret linear (conflict-free)",
    "SHFL.BFLY", "FMNMX (8 chains)", "FFMA (8 chains)", "MATCH.ANY", "F2I",
    "ATOMS.ADD ret random(512w/warp)", "STS+LDS+ATOMS random mix (3 ops)", "VIMNMX relu clamp"};

template <int T>
__global__ void __launch_bounds__(THREADS, 1) k_ops(int iters, long long* cyc, uint32_t* sink) {
    extern __shared__ uint32_t sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* h = sm + warp * 1024;
    for (int i = lane; i < 1024; i += 32) h[i] = 0;
    uint32_t idx[NELEM];
#pragma unroll
    for (int j = 0; j < NELEM; ++j) {
        uint32_t r = hash32((blockIdx.x * THREADS + threadIdx.x) * NELEM + j + 12345u);
        if (T == T_LDS_LIN || T == T_ATOM_LIN) idx[j] = (j * 32 + lane) & 1023;
        else if (T == T_ATOM_RET_2K) idx[j] = r & 511;
        else idx[j] = r & 1023;
    }
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = 1.0f + lane * 0.01f + j;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if constexpr (T == T_ATOM_NORET) {
#pragma unroll
            for (int j = 0; j < NELEM; ++j)
                asm volatile("red.shared.add.u32 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(h + idx[j])) : "memory");
        } else if constexpr (T == T_ATOM_RET || T == T_ATOM_LIN || T == T_ATOM_RET_2K) {
#pragma unroll
            for (int j = 0; j < NELEM; ++j) acc += atomicAdd(h + idx[j], 1u);
        } else if constexpr (T == T_ATOM_OR) {
#pragma unroll
            for (int j = 0; j < NELEM; ++j) acc += atomicOr(h + idx[j], 1u << (idx[j] & 31));
        } else if constexpr (T == T_STS) {
#pragma unroll
            for (int j = 0; j < NELEM; ++j)
                asm volatile("st.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(h + idx[j])), "r"(acc + j) : "memory");
        } else if constexpr (T == T_LDS || T == T_LDS_LIN) {
#pragma unroll
            for (int j = 0; j < NELEM; ++j) {
                uint32_t v;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(h + idx[j])) : "memory");
                acc += v;
            }
        } else if constexpr (T == T_STS_LDS_MIX) {
#pragma unroll
            for (int j = 0; j < NELEM; ++j) {
                uint32_t v;
                acc += atomicAdd(h + idx[j], 1u);
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(h + idx[(j + 7) & 31])) : "memory");
                acc += v;
                asm volatile("st.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(h + idx[(j + 13) & 31])), "r"(j) : "memory");
            }
        } else if constexpr (T == T_SHFL) {
#pragma unroll
            for (int j = 0; j < NELEM; ++j) idx[j] = __shfl_xor_sync(0xffffffffu, idx[j], 1 + (j & 15));
        } else if constexpr (T == T_FMNMX) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fminf(f[j], f[(j + 1) & 7] + 0.f) , f[j] = fmaxf(f[j], (float)it);
        } else if constexpr (T == T_FFMA) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], 1.0001f, f[(j + 1) & 7]);
        } else if constexpr (T == T_MATCH) {
#pragma unroll
            for (int j = 0; j < NELEM; ++j) acc += __match_any_sync(0xffffffffu, idx[j] + acc);
        } else if constexpr (T == T_F2I) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int j = 0; j < 8; ++j) { int q = __float2int_rz(f[j]); f[j] = __int_as_float(q + 0x3f800000); }
        } else if constexpr (T == T_VIMNMX) {
#pragma unroll
            for (int j = 0; j < NELEM; ++j) idx[j] = __vimin_s32_relu((int)idx[j] + it, 1023);
        }
    }
    const long long t1 = clock64();
#pragma unroll
    for (int j = 0; j < NELEM; ++j) acc += idx[j];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += __float_as_uint(f[j]);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc + h[lane];
}

template <int T>
void run_op(int sms, int iters, long long* d_cyc, uint32_t* d_sink) {
    const size_t smem = WARPS * 1024 * 4;
    CK(cudaFuncSetAttribute(k_ops<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_ops<T><<<sms, THREADS, smem>>>(10, d_cyc, d_sink);
    k_ops<T><<<sms, THREADS, smem>>>(iters, d_cyc, d_sink);
    CK(cudaDeviceSynchronize());
    long long* h = (long long*)malloc(sms * sizeof(long long));
    CK(cudaMemcpy(h, d_cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += h[i];
    avg /= sms;
    int ops = NELEM;
    if (T == T_FMNMX) ops = 64;
    if (T == T_FFMA || T == T_F2I) ops = 32;
    if (T == T_STS_LDS_MIX) ops = 96;
    const double per = avg / ((double)iters * ops * WARPS);
    printf("%-44s %8.3f cyc/warp-instr/SM  %7.2f lanes/clk/SM\n", names[T], per, 32.0 / per);
    free(h);
}

// ------------------------------------------------------------------ TMA stream
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NODES>
__global__ void __launch_bounds__(THREADS, 1)
k_stream(const __grid_constant__ CUtensorMap tmap, int R, int boxrows, long long ntiles, float* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bar = (uint64_t*)(smem + 1024 * NODES * 4);
    const int nbox = (R + boxrows - 1) / boxrows;
    auto issue = [&](long long tile) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"((uint32_t)(nbox * boxrows * NODES * 4)) : "memory");
        for (int b = 0; b < nbox; ++b)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(s32(smem + (size_t)b * boxrows * NODES * 4)), "l"(&tmap), "r"(s32(bar)), "r"((int)(tile * NODES)), "r"(b * boxrows) : "memory");
    };
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (blockIdx.x < ntiles) issue(blockIdx.x);
    }
    __syncthreads();
    uint32_t phase = 0;
    float acc = 0.f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(bar)), "r"(phase) : "memory");
        phase ^= 1;
        // every warp pulls "its" column(s) into registers like K2 does
        const float* t = (const float*)smem;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
            const int row = j * 32 + lane;
            if (NODES == 32) { const float2 v = *(const float2*)(t + row * 32 + warp * 2); acc += v.x + v.y; }
            else acc += t[row * 16 + warp];
        }
        __syncthreads();
        if (threadIdx.x == 0) { const long long nx = tile + gridDim.x; if (nx < ntiles) issue(nx); }
    }
    if (acc == 1234.5f) out[0] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int NODES>
void run_stream(EncodeFn enc, float* d, long long N, int R, int sms, float* d_out) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)R};
    cuuint64_t strides[1] = {(cuuint64_t)N * 4};
    cuuint32_t box[2] = {NODES, 256};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return; }
    const size_t smem = 1024 * NODES * 4 + 64;
    CK(cudaFuncSetAttribute(k_stream<NODES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long ntiles = N / NODES;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_stream<NODES><<<sms, THREADS, smem>>>(tm, R, 256, ntiles, d_out);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("TMA stream tile [%d rows x %d nodes] single buffer: %.3f ms  %.1f GB/s\n", R, NODES, ms, (double)N * R * 4 / ms / 1e6);
    }
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    long long* d_cyc; uint32_t* d_sink;
    CK(cudaMalloc(&d_cyc, sms * sizeof(long long)));
    CK(cudaMalloc(&d_sink, 1024));
    const int it = 2000;
    run_op<T_ATOM_NORET>(sms, it, d_cyc, d_sink);
    run_op<T_ATOM_RET>(sms, it, d_cyc, d_sink);
    run_op<T_ATOM_OR>(sms, it, d_cyc, d_sink);
    run_op<T_ATOM_RET_2K>(sms, it, d_cyc, d_sink);
    run_op<T_ATOM_LIN>(sms, it, d_cyc, d_sink);
    run_op<T_STS>(sms, it, d_cyc, d_sink);
    run_op<T_LDS>(sms, it, d_cyc, d_sink);
    run_op<T_LDS_LIN>(sms, it, d_cyc, d_sink);
    run_op<T_STS_LDS_MIX>(sms, it, d_cyc, d_sink);
    run_op<T_SHFL>(sms, it, d_cyc, d_sink);
    run_op<T_FMNMX>(sms, it, d_cyc, d_sink);
    run_op<T_FFMA>(sms, it, d_cyc, d_sink);
    run_op<T_MATCH>(sms, it / 4, d_cyc, d_sink);
    run_op<T_F2I>(sms, it, d_cyc, d_sink);
    run_op<T_VIMNMX>(sms, it, d_cyc, d_sink);

    EncodeFn enc = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr));
    const long long N = 4000000; const int R = 1000;
    float* d; CK(cudaMalloc(&d, (size_t)N * R * 4));
    CK(cudaMemset(d, 0, (size_t)N * R * 4));
    run_stream<32>(enc, d, N, R, sms, (float*)d_sink);
    run_stream<16>(enc, d, N, R, sms, (float*)d_sink);
    return 0;
}
