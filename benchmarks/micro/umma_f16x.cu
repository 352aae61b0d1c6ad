// Error measurement for the mixed-precision split of the retrieval filter (retrieval_tc.cuh, round 2):
//   D = S * (A_hi B_hi)            kind::tf32, K = 32   (hi = the 11 leading bits of the float32 value, exact in TF32)
//     + [S A_lo | A_hi] [B_hi | S B_lo]^T   kind::f16, K = 64: both cross terms as ONE fp16 GEMM (S = 2^11 keeps the
//                                            residuals in fp16's normal range; hi is exact in fp16 above 2^-14)
// into the same float32 TMEM accumulator: 8 + 8 MMAs per 256 x 128 tile instead of the 24 of the three TF32 passes.
// SPLIT = 1 is the old three-pass scheme, SPLIT = 2 the mixed one, SPLIT = 3 the final all-fp16 one: ONE operand format
// [S hi | S lo] (64 halves = one 128-byte row) for both sides, the three terms are K slices of the same rows
// (A[32..63] x B[0..31], A[0..31] x B[32..63], A[0..31] x B[0..31]; accumulator = S^2 * score), 6 MMAs of K = 16 per
// 128 x 128 tile; the B tile is written to global memory in its swizzled image and fetched with cp.async.bulk, as the
// kernel does.  All are compared with float64 on the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_f16x umma_f16x.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int M = 128, N = 128, K = 32;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, int max_spins = 1 << 22)
{
    for (int i = 0; i < max_spins; i++) {
        uint32_t ok;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

// K-major, SWIZZLE_128B: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_desc(const void *tile)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(tile) >> 4) & 0x3fff);       // start address
    d |= (uint64_t)0 << 16;                                 // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                       // stride byte offset: 8 rows
    d |= (uint64_t)1 << 46;                                 // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                                 // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ uint32_t swz(int row, int col)   // byte offset of float (row, col) in a [rows][32] tile
{
    const int chunk = col >> 2;
    return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4) + (col & 3) * 4);
}

// SPLIT: operands are split into tf32-exact high parts and residuals in shared memory and the product is
// accumulated as A_hi B_hi + A_hi B_lo + A_lo B_hi (a K = 96 chain): float32-grade scores from tf32 tensor cores.
constexpr float XS = 2048.f;
__device__ __forceinline__ uint32_t swz_h(int row, int col)   // byte offset of half (row, col) in a [rows][64] tile
{
    return (uint32_t)(row * 128 + (((col >> 3) ^ (row & 7)) << 4) + (col & 7) * 2);
}

template <int SPLIT>
__global__ void __launch_bounds__(128) k(const float *A, const float *B, float *C, int *status, unsigned char *scratch)
{
    extern __shared__ __align__(1024) unsigned char raw[];
    unsigned char *base = (unsigned char *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    unsigned char *sA = base, *sB = base + M * 128, *sAl = base + (M + N) * 128, *sBl = sAl + M * 128;
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < M * K; e += 128) {
        const float v = A[e], hi = SPLIT ? __uint_as_float(__float_as_uint(v) & 0xffffe000u) : v;
        const int r = e / K, c = e % K;
        *(float *)(sA + swz(r, c)) = SPLIT == 2 ? hi * XS : hi;
        if (SPLIT == 1) *(float *)(sAl + swz(r, c)) = v - hi;
        if (SPLIT == 2) {
            *(__half *)(sAl + swz_h(r, c)) = __float2half_rn((v - hi) * XS);
            *(__half *)(sAl + swz_h(r, 32 + c)) = __float2half_rn(hi);
        }
        if (SPLIT == 3) {
            *(__half *)(sAl + swz_h(r, c)) = __float2half_rn(hi * XS);
            *(__half *)(sAl + swz_h(r, 32 + c)) = __float2half_rn((v - hi) * XS);
        }
    }
    for (int e = tid; e < N * K; e += 128) {
        const float v = B[e], hi = SPLIT ? __uint_as_float(__float_as_uint(v) & 0xffffe000u) : v;
        const int r = e / K, c = e % K;
        *(float *)(sB + swz(r, c)) = hi;
        if (SPLIT == 1) *(float *)(sBl + swz(r, c)) = v - hi;
        if (SPLIT == 2) {
            *(__half *)(sBl + swz_h(r, c)) = __float2half_rn(hi);
            *(__half *)(sBl + swz_h(r, 32 + c)) = __float2half_rn((v - hi) * XS);
        }
        if (SPLIT == 3) {                                      // image in global memory; the bulk copy below brings it in
            *(__half *)(scratch + swz_h(r, c)) = __float2half_rn(hi * XS);
            *(__half *)(scratch + swz_h(r, 32 + c)) = __float2half_rn((v - hi) * XS);
        }
    }
    __shared__ uint64_t bar_b;
    if (SPLIT == 3) {
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");       // generic-proxy global writes -> the bulk copy's async proxy
        __syncthreads();
        if (tid == 0) {
            mbar_init(&bar_b, 1);
            asm volatile("fence.mbarrier_init.release.cluster;");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_b)), "r"(N * 128) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sBl)), "l"(scratch), "r"(N * 128), "r"(smem_u32(&bar_b)) : "memory");
        }
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");          // generic-proxy tile writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = slot;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        if (SPLIT == 3) {
            if (!mbar_wait(&bar_b, 0)) *status = 2;
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint32_t idesc16 = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // f16 x f16 -> f32
            const uint64_t ad = make_desc(sAl), bd = make_desc(sBl);
            const uint64_t as[6] = {ad + 4, ad + 6, ad, ad + 2, ad, ad + 2}, bs[6] = {bd, bd + 2, bd + 4, bd + 6, bd, bd + 2};
            for (int i = 0; i < 6; i++) {
                const uint32_t acc = i > 0;
                asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
                             ::"r"(tmem), "l"(as[i]), "l"(bs[i]), "r"(idesc16), "r"(acc));
            }
        } else if (SPLIT == 2) {
            const uint32_t idesc16 = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // f16 x f16 -> f32
            const uint64_t ax = make_desc(sAl), bx = make_desc(sBl), ah = make_desc(sA), bh = make_desc(sB);
            for (int kk = 0; kk < 4; kk++) {                   // small terms first
                const uint32_t acc = kk > 0;
                asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
                             ::"r"(tmem), "l"(ax + 2 * kk), "l"(bx + 2 * kk), "r"(idesc16), "r"(acc));
            }
            for (int kk = 0; kk < 4; kk++)
                asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
                             ::"r"(tmem), "l"(ah + 2 * kk), "l"(bh + 2 * kk), "r"(idesc), "r"(1u));
        } else {
        const uint64_t da[3] = {make_desc(sAl), make_desc(sA), make_desc(sA)};      // small terms first
        const uint64_t db[3] = {make_desc(sB), make_desc(sBl), make_desc(sB)};
        for (int term = SPLIT ? 0 : 2; term < 3; term++)
            for (int kk = 0; kk < K / 8; kk++) {
                const uint64_t a = da[term] + (uint64_t)(kk * 32 >> 4), b = db[term] + (uint64_t)(kk * 32 >> 4);
                const uint32_t acc = !(kk == 0 && term == (SPLIT ? 0 : 2));
                asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
                             ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc));
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    }
    const bool ok = mbar_wait(&bar, 0);
    if (!ok) { if (tid == 0) *status = 1; }
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (ok) {
        const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16);
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                         "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr + c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 32; j++) C[(size_t)(32 * warp + lane) * N + c0 + j] = __uint_as_float(v[j]) * (SPLIT == 3 ? 1.f / (XS * XS) : SPLIT == 2 ? 1.f / XS : 1.f);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
}

static int run(int split, const std::vector<float> &A, const std::vector<float> &B, double tol)
{
    std::vector<float> C(M * N, -1.f);
    float *dA, *dB, *dC;
    int *dS, st = 0;
    unsigned char *dX;
    cudaMalloc(&dX, N * 128);
    cudaMemset(dX, 0, N * 128);
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dC, C.size() * 4); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dS, 0, 4);
    const size_t smem = 2 * (M + N) * 128 + 1024;
    if (split == 3) {
        cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<3><<<1, 128, smem>>>(dA, dB, dC, dS, dX);
    } else if (split == 2) {
        cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<2><<<1, 128, smem>>>(dA, dB, dC, dS, dX);
    } else if (split == 1) {
        cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<1><<<1, 128, smem>>>(dA, dB, dC, dS, dX);
    } else {
        cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<0><<<1, 128, smem>>>(dA, dB, dC, dS, dX);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("split=%d launch: %s\n", split, cudaGetErrorString(e));
    cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    printf("status (1 = mbarrier timeout, 2 = bulk copy timeout): %d\n", st);
    int bad = 0;
    double maxerr = 0;
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            double ref = 0;
            for (int c = 0; c < K; c++) ref += (double)A[i * K + c] * B[j * K + c];
            const double err = fabs(ref - C[i * N + j]);
            if (err > maxerr) maxerr = err;
            if (err > tol) { if (bad < 5) printf("mismatch C[%d][%d] = %g want %g\n", i, j, C[i * N + j], ref); bad++; }
        }
    printf("mismatches: %d of %d, max err %g (tolerance %g)\n", bad, M * N, maxerr, tol);
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dS); cudaFree(dX);
    return bad != 0 || st != 0 || e != cudaSuccess;
}

int main()
{
    std::vector<float> A(M * K), B(N * K);
    srand(1);
    for (auto &v : A) v = (float)(rand() % 33 - 16) / 8.f;      // exactly representable in tf32 and fp16
    for (auto &v : B) v = (float)(rand() % 33 - 16) / 8.f;
    int rc = run(0, A, B, 1e-5);
    rc |= run(2, A, B, 1e-5);
    rc |= run(3, A, B, 1e-5);
    // unit vectors of dimension 26 (zero padded to 32), several magnitudes of clustering
    for (int trial = 0; trial < 3; trial++) {
        const double spread = trial == 0 ? 1.0 : (trial == 1 ? 0.05 : 0.001);
        auto fill = [&](std::vector<float> &X, int rows) {
            for (int r = 0; r < rows; r++) {
                double nrm = 0, t[32];
                for (int c = 0; c < 26; c++) { t[c] = 1.0 + spread * ((rand() % 20001 - 10000) / 10000.0) * (c + 1); nrm += t[c] * t[c]; }
                for (int c = 0; c < 32; c++) X[r * K + c] = c < 26 ? (float)(t[c] / sqrt(nrm)) : 0.f;
            }
        };
        fill(A, M);
        fill(B, N);
        printf("-- unit vectors, spread %g\n", spread);
        rc |= run(1, A, B, 2e-6);
        rc |= run(2, A, B, 2e-6);
        rc |= run(3, A, B, 2e-6);
    }
    // wide dynamic range: a few large components and many tiny ones (down to 1e-9 of the norm): the small high parts and
    // scaled residuals fall into fp16's subnormal range (or flush to zero) -- the bound in retrieval_tc.cuh covers both
    for (int trial = 0; trial < 3; trial++) {
        auto fill = [&](std::vector<float> &X, int rows) {
            for (int r = 0; r < rows; r++) {
                double nrm = 0, t[32];
                for (int c = 0; c < 26; c++) {
                    const double mag = pow(10.0, -(double)(rand() % (3 + 3 * trial + 1)));       // 1 ... 1e-3 / 1e-6 / 1e-9
                    t[c] = mag * ((rand() % 20001 - 10000) / 10000.0);
                    nrm += t[c] * t[c];
                }
                if (nrm == 0) { t[0] = 1; nrm = 1; }
                for (int c = 0; c < 32; c++) X[r * K + c] = c < 26 ? (float)(t[c] / sqrt(nrm)) : 0.f;
            }
        };
        fill(A, M);
        fill(B, N);
        printf("-- wide dynamic range, components down to 1e-%d of the largest\n", 3 + 3 * trial);
        rc |= run(1, A, B, 2e-6);
        rc |= run(2, A, B, 2e-6);
        rc |= run(3, A, B, 2e-6);
    }
    // adversarial mantissas: every value has its 13 low bits set (largest possible residual, all of one sign)
    for (int trial = 0; trial < 2; trial++) {
        auto fill = [&](std::vector<float> &X, int rows) {
            for (int r = 0; r < rows; r++) {
                double nrm = 0, t[32];
                for (int c = 0; c < 26; c++) { t[c] = 1.0 + (trial ? 0.3 : 0.001) * ((rand() % 20001 - 10000) / 10000.0); nrm += t[c] * t[c]; }
                for (int c = 0; c < 32; c++) {
                    float v = c < 26 ? (float)(t[c] / sqrt(nrm)) : 0.f;
                    uint32_t b;
                    memcpy(&b, &v, 4);
                    if (c < 26) b |= 0x1fffu;
                    memcpy(&v, &b, 4);
                    X[r * K + c] = v;
                }
            }
        };
        fill(A, M);
        fill(B, N);
        printf("-- adversarial low mantissa bits, trial %d\n", trial);
        rc |= run(1, A, B, 8e-6);
        rc |= run(2, A, B, 8e-6);
        rc |= run(3, A, B, 8e-6);
    }
    printf("rc=%d\n", rc);
    return rc;
}
