// gsb_synth.cu -- seeded synthetic realization stacks, device side
//
// Bit-identical twin of gsimcli_b200/synth.py::values16 (64-bit integer
// arithmetic only; see that file for the value model).  One thread produces
// 4 adjacent nodes of one realization and writes them with a 16-byte store.
#include "gsb_common.cuh"

namespace {

__device__ __forceinline__ uint64_t splitmix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ long long tri_centered(long long i, long long d) {
    const long long a = (i * 2048) / d;
    return (a < 1024 ? a : 2047 - a) - 512;
}

__device__ __forceinline__ float synth_value(uint64_t seed, long long r, long long node,
                                             long long dx, long long dy, float nodata) {
    const long long x = node % dx;
    const long long y = (node / dx) % dy;
    const long long z = node / (dx * dy);
    const long long cx = tri_centered(x, dx), cy = tri_centered(y, dy);
    const long long field = (((cx * cy + (1ll << 18)) * 4800) >> 18) - 4800;
    const long long mu = 12800 + field + 80 * (z % 100);   // trend restarts every 100 layers
    const uint64_t key = (seed * 0x94D049BB133111EBull +
                          static_cast<uint64_t>(r) * 0x9E3779B97F4A7C15ull) ^
                         (static_cast<uint64_t>(node) * 0xBF58476D1CE4E5B9ull);
    const uint64_t h = splitmix(splitmix(key));
    const long long u1 = static_cast<long long>(h & 0xFFFF);
    const long long u2 = static_cast<long long>((h >> 16) & 0xFFFF);
    const long long u3 = static_cast<long long>((h >> 32) & 0xFFFF);
    const long long sel = static_cast<long long>((h >> 48) & 0xFFFF);
    const long long noise = ((u1 * u2 * 9600) >> 32) + ((u3 * 1600) >> 16) - 3200;
    long long v = mu + noise;
    v = v < 16 * 350 ? 16 * 350 : (v > 16 * 1400 ? 16 * 1400 : v);
    return sel < 66 ? nodata : static_cast<float>(v) * 0.0625f;
}

__global__ void synth_kernel(float* data, long long pitch, long long R, long long N,
                             uint64_t seed, long long dx, long long dy, long long node0,
                             float nodata) {
    const long long quads = (N + 3) / 4;
    const long long total = R * quads;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = e / quads, qd = e % quads;
        const long long n = qd * 4;
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            v[k] = (n + k < N) ? synth_value(seed, r, node0 + n + k, dx, dy, nodata) : 0.f;
        float* dst = data + r * pitch + n;
        if (n + 3 < N) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        else
            for (int k = 0; k < 4 && n + k < N; ++k) dst[k] = v[k];
    }
}

}  // namespace

int gsb_launch_synth(gsb_stack* s, uint64_t seed, int64_t dx, int64_t dy, int64_t dz,
                     int64_t node0, float nodata, cudaStream_t st) {
    (void)dz;
    const long long total = s->R * ((s->N + 3) / 4);
    if (total <= 0) return GSB_OK;
    long long blocks = (total + 255) / 256;
    const long long cap = static_cast<long long>(s->sm_count) * 32;
    if (blocks > cap) blocks = cap;
    synth_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(s->data, s->pitch, s->R, s->N, seed,
                                                         dx, dy, node0, nodata);
    GSB_CUDA(cudaGetLastError());
    return GSB_OK;
}
