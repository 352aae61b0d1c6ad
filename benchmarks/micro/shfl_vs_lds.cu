// Microbenchmark: does SHFL share the L1 data pipe with shared-memory accesses on sm_100a?
// Three kernels with the same loop count: (a) LDS.128 only, (b) SHFL only, (c) both interleaved.
// If (c) ~ max(a, b) the two overlap; if (c) ~ a + b they serialise on one pipe.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

__global__ void k_lds(float4 *out, int n)
{
    __shared__ float4 buf[256 * 4];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    int idx = threadIdx.x;
    for (int it = 0; it < ITERS; it++) {
        float4 v = buf[idx];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        idx = (idx + 33) & 1023;
    }
    if (n == 12345) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void k_shfl(float4 *out, int n)
{
    float a = threadIdx.x, b = a + 1, c = a + 2, d = a + 3;
    for (int it = 0; it < ITERS; it++) {
        a += __shfl_xor_sync(0xffffffffu, b, 1);
        b += __shfl_xor_sync(0xffffffffu, c, 2);
        c += __shfl_xor_sync(0xffffffffu, d, 4);
        d += __shfl_xor_sync(0xffffffffu, a, 8);
    }
    if (n == 12345) out[blockIdx.x * blockDim.x + threadIdx.x] = make_float4(a, b, c, d);
}

__global__ void k_both(float4 *out, int n)
{
    __shared__ float4 buf[256 * 4];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    float a = threadIdx.x, b = a + 1, c = a + 2, d = a + 3;
    int idx = threadIdx.x;
    for (int it = 0; it < ITERS; it++) {
        float4 v = buf[idx];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        idx = (idx + 33) & 1023;
        a += __shfl_xor_sync(0xffffffffu, b, 1);
        b += __shfl_xor_sync(0xffffffffu, c, 2);
        c += __shfl_xor_sync(0xffffffffu, d, 4);
        d += __shfl_xor_sync(0xffffffffu, a, 8);
    }
    if (n == 12345) out[blockIdx.x * blockDim.x + threadIdx.x] = make_float4(a + acc.x, b + acc.y, c + acc.z, d + acc.w);
}

template <typename K>
float run(K k, const char *name, float4 *out)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<148 * 2, 512>>>(out, 0);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) k<<<148 * 2, 512>>>(out, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    // per SM: 2 CTAs x 16 warps, ITERS iterations
    const double warp_iters_per_sm = 2.0 * 16 * ITERS;
    printf("%-8s %.3f ms  -> %.2f cycles per warp-iteration per SM at 1.965 GHz\n", name, ms, ms * 1e-3 * 1.965e9 / warp_iters_per_sm);
    return ms;
}

int main()
{
    float4 *out; cudaMalloc(&out, 148 * 2 * 512 * sizeof(float4));
    run(k_lds, "lds128", out);     // 1 LDS.128 (4 wavefronts) per iteration
    run(k_shfl, "shfl x4", out);   // 4 SHFL per iteration
    run(k_both, "both", out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
