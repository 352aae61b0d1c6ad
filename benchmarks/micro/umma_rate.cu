// tcgen05.mma issue-rate calibration: one CTA per SM issues REPS x 12 MMAs (M 128, N, K 8 tf32 / K 16 bf16) on the
// same shared-memory operands and waits for the commit; prints cycles per MMA and the implied dense TFLOP/s.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(const void *tile)
{
    return (uint64_t)((smem_u32(tile) >> 4) & 0x3fff) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

template <int N, bool BF16>
__global__ void __launch_bounds__(128) k(int reps, long long *cycles)
{
    extern __shared__ __align__(1024) unsigned char raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < (128 + N) * 128 / 4; e += 128) ((float *)raw)[e] = 0.f;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = slot;
    if (tid == 0) {
        const uint32_t fmt = BF16 ? 1u : 2u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da = make_desc(raw), db = make_desc(raw + 128 * 128);
        const long long t0 = clock64();
        for (int r = 0; r < reps; r++) {
#pragma unroll
            for (int i = 0; i < 12; i++) {
                const uint64_t a = da + 2 * (i & 3), b = db + 2 * (i & 3);
                const uint32_t acc = 1;
                if (BF16)
                    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
                                 ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc));
                else
                    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
                                 ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc));
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        if (blockIdx.x == 0) *cycles = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

template <int N, bool BF16>
void run(const char *name)
{
    long long *d, h = 0;
    cudaMalloc(&d, 8);
    const int reps = 2000;
    const size_t smem = (128 + N) * 128;
    cudaFuncSetAttribute(k<N, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<N, BF16><<<148, 128, smem>>>(10, d);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<N, BF16><<<148, 128, smem>>>(reps, d);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const double kk = BF16 ? 16 : 8;
    const double flop = 148.0 * reps * 12 * 2.0 * 128 * N * kk;
    printf("%-28s %s  %.1f cycles/MMA  %.3f ms  %.1f TFLOP/s\n", name, cudaGetErrorString(e), (double)h / (reps * 12.0), ms, flop / ms / 1e9);
    cudaFree(d);
}

int main()
{
    run<128, false>("tf32 M128 N128 K8");
    run<256, false>("tf32 M128 N256 K8");
    run<128, true>("bf16 M128 N128 K16");
    run<256, true>("bf16 M128 N256 K16");
    return 0;
}
