// Microbenchmark: can TMEM serve as a per-lane constant store whose reads (tcgen05.ld) do not compete with
// shared-memory traffic on the L1 data pipe?  (a) LDS.128 only  (b) tcgen05.ld x4 only  (c) both.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

__device__ __forceinline__ uint32_t tmem_alloc(uint32_t *slot, int warp)
{
    if (warp == 0) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(slot);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(sa));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    return *slot;
}
__device__ __forceinline__ void tmem_free(uint32_t base, int warp)
{
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(base));
}
__device__ __forceinline__ void tmem_st4(uint32_t addr, float a, float b, float c, float d)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d));
}
__device__ __forceinline__ void tmem_ld4(uint32_t addr, float &a, float &b, float &c, float &d)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(addr));
}

template <int MODE>
__global__ void k(float *out, int n, int *bad)
{
    __shared__ float4 buf[1024];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_float4(i, 1, 2, 3);
    const uint32_t base = tmem_alloc(&slot, warp);
    const uint32_t mine = base + ((uint32_t)(32 * (warp & 3)) << 16);
    // every warp of a lane quarter writes the same values: value = 1000 * column + lane
    for (int c = 0; c < 64; c += 4)
        tmem_st4(mine + c, 1000.f * c + lane, 1000.f * (c + 1) + lane, 1000.f * (c + 2) + lane, 1000.f * (c + 3) + lane);
    asm volatile("tcgen05.wait::st.sync.aligned;");
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    float t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    int idx = threadIdx.x, col = 0;
    for (int it = 0; it < ITERS; it++) {
        if (MODE != 1) {
            const float4 v = buf[idx];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            idx = (idx + 33) & 1023;
        }
        if (MODE != 0) {
            float a, b, c, d;
            tmem_ld4(mine + col, a, b, c, d);
            asm volatile("tcgen05.wait::ld.sync.aligned;");
            if (a != 1000.f * col + lane || d != 1000.f * (col + 3) + lane) atomicAdd(bad, 1);
            t0 += a; t1 += b; t2 += c; t3 += d;
            col = (col + 4) & 63;
        }
    }
    if (n == 12345) out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w + t0 + t1 + t2 + t3;
    tmem_free(base, warp);
}

template <int MODE>
void run(const char *name, float *out, int *bad)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 2, 512>>>(out, 0, bad);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) k<MODE><<<148 * 2, 512>>>(out, 0, bad);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    printf("%-10s %.3f ms -> %.2f cycles per warp-iteration per SM (1.965 GHz)\n", name, ms, ms * 1e-3 * 1.965e9 / (2.0 * 16 * ITERS));
}

int main()
{
    float *out; int *bad, hbad = -1;
    cudaMalloc(&out, 148 * 2 * 512 * sizeof(float));
    cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
    run<0>("lds128", out, bad);
    run<1>("ldtm x4", out, bad);
    run<2>("both", out, bad);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&hbad, bad, 4, cudaMemcpyDeviceToHost);
    printf("status %s, mismatches %d\n", cudaGetErrorString(e), hbad);
    return 0;
}
