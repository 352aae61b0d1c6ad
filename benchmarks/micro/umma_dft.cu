// umma_dft.cu -- prototype of ONE FFT stage as a tensor-core GEMM (512 = 32 x 16, split TF32).
//
// What it measures.  The first stage of a 512-point complex FFT (1024 real samples packed as z[n] = x[2n] + i x[2n+1],
// n = 16 n1 + n2):   Y[k1, n2] = sum_{n1 < 32} z[16 n1 + n2] W_32^{n1 k1}
// as the GEMM  D[(frame, n2), (k1, re|im)] = A[(frame, n2), (n1, re|im)] . B[(k1, re|im), (n1, re|im)]^T  with M = 128 rows
// (8 frames x 16 n2), N = 64, K = 64, issued as tcgen05.mma kind::tf32 in three passes (A_lo B_hi + A_hi B_lo +
// A_hi B_hi: plain TF32 is 8e-4, outside the 1e-4 tolerance of the features).  Roles, as in retrieval_tc.cuh:
//   warps 0-3  readers: tcgen05.ld of their TMEM lane (64 columns), |Y|^2 accumulated in registers (what a next
//              stage living in the same thread would consume) or, with verify, the rows written to global memory
//   warps 4-7  loaders: coalesced 8-byte global loads of the strided samples, hi / lo split, 128-bit stores into the
//              K-major SWIZZLE_128B operand tiles (operands written exactly once)
//   warp 8     one lane issues the 24 MMAs of a tile (M 128, N 64, K 8) and commits to the mbarriers
// Modes: 0 full pipeline, 1 MMA + readers on resident operands (no loader work), 2 MMA only, 3 loaders only,
//        4 readers only.  Each prints cycles per frame per SM -- next to the 395 cycles per frame that the whole
//        warp-autonomous FP32 kernel (feat_warp8) spends on its three FFT passes, window, mel, log and DCT together.
// verify: max |D - D_ref| / ||row|| against a float64 DFT on the host.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_dft umma_dft.cu -cudart shared
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>

constexpr int ROWS = 128, KDIM = 64, NOUT = 64, FRAMES_PER_TILE = 8, SAMPLES = 1024;
constexpr int STAGES = 2;
constexpr int TILE_BYTES = ROWS * 128;                       // one 32-float-wide K half of the A tile
constexpr int A_STAGE_BYTES = 4 * TILE_BYTES;                // hi K0, hi K1, lo K0, lo K1
constexpr int B_TILE_BYTES = NOUT * 128;                     // 64 rows x 128 B
constexpr int THREADS = 288;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(const void *tile)
{
    return (uint64_t)((smem_u32(tile) >> 4) & 0x3fff) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    for (uint32_t spin = 0;; spin++) {
        uint32_t ok;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
        if (spin > (1u << 28)) __trap();
    }
}
__device__ __forceinline__ void mma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float4 hi4(float4 v)
{
    return make_float4(__uint_as_float(__float_as_uint(v.x) & 0xffffe000u), __uint_as_float(__float_as_uint(v.y) & 0xffffe000u),
                       __uint_as_float(__float_as_uint(v.z) & 0xffffe000u), __uint_as_float(__float_as_uint(v.w) & 0xffffe000u));
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct Params {
    const float *x;          // [n_frames][1024]
    const float *bmat;       // [2 (hi, lo)][2 (K half)][64 rows][32] DFT matrix, split on the host
    float *out;              // verify: [n_frames * 16][64]
    double *sums;            // perf: one checksum per CTA
    long long *cyc;          // [148][4]: loader, mma, reader, total cycles of each CTA
    int n_tiles, mode, verify;
};

__global__ void __launch_bounds__(THREADS, 1) dft_stage_kernel(const Params p)
{
    extern __shared__ __align__(1024) unsigned char raw[];
    unsigned char *a_tiles = raw;                                     // [STAGES][4][ROWS x 128 B]
    unsigned char *b_tiles = raw + STAGES * A_STAGE_BYTES;            // [hi, lo][K half][64 x 128 B]
    uint64_t *bars = reinterpret_cast<uint64_t *>(b_tiles + 4 * B_TILE_BYTES);
    uint64_t *full_a = bars, *empty_a = bars + STAGES, *acc_full = bars + 2 * STAGES, *acc_empty = acc_full + 2;
    uint32_t *slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool use_loader = p.mode == 0 || p.mode == 3, use_mma = p.mode <= 2, use_reader = p.mode == 0 || p.mode == 1 || p.mode == 4;

    // B tiles: K-major rows of 128 bytes under the 128-byte swizzle
    for (int e = tid; e < 4 * NOUT * 8; e += THREADS) {
        const int t = e / (NOUT * 8), row = (e / 8) % NOUT, chunk = e % 8;
        const float4 v = *reinterpret_cast<const float4 *>(p.bmat + ((size_t)t * NOUT + row) * 32 + chunk * 4);
        *reinterpret_cast<float4 *>(b_tiles + t * B_TILE_BYTES + row * 128 + ((chunk ^ (row & 7)) << 4)) = v;
    }
    for (int e = tid; e < STAGES * A_STAGE_BYTES / 16; e += THREADS) reinterpret_cast<float4 *>(a_tiles)[e] = make_float4(0.5f, -0.25f, 0.125f, 1.f);
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full_a[s], 128); mbar_init(&empty_a[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *slot;
    const long long t_start = clock64();
    long long t_role = 0;

    if (warp < 4) {
        // ===== readers: one GEMM row (frame, n2) per thread =====
        if (use_reader) {
            float acc = 0.f;
            for (int t = 0; t < p.n_tiles; t++) {
                const int buf = t & 1;
                const int64_t tile = (int64_t)blockIdx.x * p.n_tiles + t;
                if (use_mma) mbar_wait(&acc_full[buf], (uint32_t)((t >> 1) & 1));
                __syncwarp();
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * NOUT);
                uint32_t v[32];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    ld32(taddr + h * 32, v);
                    if (p.verify) {
                        float4 *dst = reinterpret_cast<float4 *>(p.out + ((size_t)tile * ROWS + tid) * NOUT + h * 32);
#pragma unroll
                        for (int j = 0; j < 8; j++)
                            dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j++) acc = fmaf(__uint_as_float(v[j]), __uint_as_float(v[j]), acc);
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;");
                __syncwarp();
                if (use_mma && lane == 0) mbar_arrive(&acc_empty[buf]);
            }
            if (!p.verify) atomicAdd(p.sums + blockIdx.x, (double)acc);
        }
        t_role = clock64() - t_start;
        if (tid == 0) p.cyc[blockIdx.x * 4 + 2] = t_role;
    } else if (warp < 8) {
        // ===== loaders: thread = GEMM row (frame f of the tile, n2); K runs over n1 (re, im) =====
        if (use_loader) {
            const int r = tid - 128, f = r >> 4, n2 = r & 15;
            for (int t = 0; t < p.n_tiles; t++) {
                const int s = t % STAGES;
                const int64_t tile = (int64_t)blockIdx.x * p.n_tiles + t;
                const float *src = p.x + ((size_t)tile * FRAMES_PER_TILE + f) * SAMPLES + 2 * n2;
                float2 v[32];
#pragma unroll
                for (int n1 = 0; n1 < 32; n1++) v[n1] = __ldg(reinterpret_cast<const float2 *>(src + 32 * n1));   // 16 lanes = one 128-byte line
                if (use_mma) mbar_wait(&empty_a[s], (uint32_t)(((t / STAGES) & 1) ^ 1));
                unsigned char *st = a_tiles + s * A_STAGE_BYTES;
#pragma unroll
                for (int c = 0; c < 16; c++) {                         // chunk c = K elements 4c .. 4c+3 = n1 = 2c, 2c+1
                    const float4 val = make_float4(v[2 * c].x, v[2 * c].y, v[2 * c + 1].x, v[2 * c + 1].y);
                    const float4 h = hi4(val);
                    const uint32_t off = (uint32_t)((c >> 3) * TILE_BYTES + r * 128 + (((c & 7) ^ (r & 7)) << 4));
                    *reinterpret_cast<float4 *>(st + off) = h;
                    *reinterpret_cast<float4 *>(st + 2 * TILE_BYTES + off) = make_float4(val.x - h.x, val.y - h.y, val.z - h.z, val.w - h.w);
                }
                asm volatile("fence.proxy.async.shared::cta;");
                if (use_mma) mbar_arrive(&full_a[s]);
            }
        }
        t_role = clock64() - t_start;
        if (tid == 128) p.cyc[blockIdx.x * 4 + 0] = t_role;
    } else if (lane == 0) {
        // ===== tensor-core issue =====
        if (use_mma) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NOUT >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
            for (int t = 0; t < p.n_tiles; t++) {
                const int s = t % STAGES, buf = t & 1;
                if (use_reader) mbar_wait(&acc_empty[buf], (uint32_t)(((t >> 1) & 1) ^ 1));
                if (use_loader) mbar_wait(&full_a[s], (uint32_t)((t / STAGES) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;");
                const unsigned char *st = a_tiles + s * A_STAGE_BYTES;
                const uint32_t d = tmem + (uint32_t)(buf * NOUT);
                // passes: A_lo B_hi, A_hi B_lo, A_hi B_hi (small terms first); K = 64 = 2 tiles x 4 K-steps of 8
#pragma unroll
                for (int pass = 0; pass < 3; pass++) {
                    const int a_lo = pass == 0, b_lo = pass == 1;
#pragma unroll
                    for (int kt = 0; kt < 2; kt++) {
                        const uint64_t da = make_desc(st + (a_lo ? 2 : 0) * TILE_BYTES + kt * TILE_BYTES);
                        const uint64_t db = make_desc(b_tiles + ((b_lo ? 2 : 0) + kt) * B_TILE_BYTES);
#pragma unroll
                        for (int kk = 0; kk < 4; kk++) mma_tf32(d, da + 2 * kk, db + 2 * kk, idesc, (pass | kt | kk) != 0);
                    }
                }
                if (use_loader) commit(&empty_a[s]);
                commit(&acc_full[buf]);
                if (!use_reader && (t & 15) == 15) mbar_wait(&acc_full[buf], (uint32_t)((t >> 1) & 1));   // keep the queue bounded
            }
            if (!use_reader) mbar_wait(&acc_full[(p.n_tiles - 1) & 1], (uint32_t)(((p.n_tiles - 1) >> 1) & 1));
        }
        t_role = clock64() - t_start;
        p.cyc[blockIdx.x * 4 + 1] = t_role;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (tid == 0) p.cyc[blockIdx.x * 4 + 3] = clock64() - t_start;
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

int main(int argc, char **argv)
{
    const int n_ctas = 148, tiles = argc > 1 ? atoi(argv[1]) : 256;
    const int64_t n_frames = (int64_t)n_ctas * tiles * FRAMES_PER_TILE;
    // DFT-32 matrix as a real 64 x 64: out (k1, re) = sum re cos + im sin, out (k1, im) = sum im cos - re sin
    std::vector<float> bm(2 * 2 * NOUT * 32);
    std::vector<double> bd((size_t)NOUT * KDIM);
    for (int k1 = 0; k1 < 32; k1++)
        for (int n1 = 0; n1 < 32; n1++) {
            const double th = 2.0 * M_PI * (double)((n1 * k1) % 32) / 32.0, c = std::cos(th), s = std::sin(th);
            bd[(size_t)(2 * k1) * KDIM + 2 * n1] = c;      bd[(size_t)(2 * k1) * KDIM + 2 * n1 + 1] = s;
            bd[(size_t)(2 * k1 + 1) * KDIM + 2 * n1] = -s; bd[(size_t)(2 * k1 + 1) * KDIM + 2 * n1 + 1] = c;
        }
    for (int row = 0; row < NOUT; row++)
        for (int k = 0; k < KDIM; k++) {
            const float v = (float)bd[(size_t)row * KDIM + k];
            uint32_t u;
            memcpy(&u, &v, 4);
            u &= 0xffffe000u;
            float h;
            memcpy(&h, &u, 4);
            const int kt = k / 32, kk = k % 32;
            bm[((size_t)(0 * 2 + kt) * NOUT + row) * 32 + kk] = h;
            bm[((size_t)(1 * 2 + kt) * NOUT + row) * 32 + kk] = v - h;
        }
    std::vector<float> hx((size_t)8 * FRAMES_PER_TILE * SAMPLES);         // verification set: the first 8 tiles
    uint32_t seed = 12345u;
    for (auto &v : hx) { seed = seed * 1664525u + 1013904223u; v = ((seed >> 8) & 0xffff) / 32768.f - 1.f; }
    for (int i = 0; i < 1024; i++) hx[i] = (i % 7 == 0) ? 1.f : 1e-4f * hx[i];      // one frame with a 80 dB dynamic range
    float *dx, *dbm, *dout;
    double *dsum;
    long long *dcyc;
    CK(cudaMalloc(&dx, (size_t)n_frames * SAMPLES * 4));
    CK(cudaMalloc(&dbm, bm.size() * 4));
    CK(cudaMalloc(&dout, (size_t)8 * ROWS * NOUT * 4));
    CK(cudaMalloc(&dsum, n_ctas * 8));
    CK(cudaMalloc(&dcyc, n_ctas * 4 * 8));
    CK(cudaMemset(dx, 0, (size_t)n_frames * SAMPLES * 4));
    for (int64_t off = 0; off < n_frames * SAMPLES; off += (int64_t)hx.size())                       // tile the random block
        CK(cudaMemcpy(dx + off, hx.data(), std::min<int64_t>(hx.size(), n_frames * SAMPLES - off) * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dbm, bm.data(), bm.size() * 4, cudaMemcpyHostToDevice));
    const size_t smem = STAGES * A_STAGE_BYTES + 4 * B_TILE_BYTES + 256;
    CK(cudaFuncSetAttribute(dft_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    // ---- parity: 8 tiles on one CTA against the float64 DFT ----
    Params p{dx, dbm, dout, dsum, dcyc, 8, 0, 1};
    dft_stage_kernel<<<1, THREADS, smem>>>(p);
    CK(cudaDeviceSynchronize());
    std::vector<float> ho((size_t)8 * ROWS * NOUT);
    CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0.0, worst_plain = 0.0;
    for (int row = 0; row < 8 * ROWS; row++) {
        const int f = row / 16, n2 = row % 16;
        double nrm = 0.0;
        for (int n1 = 0; n1 < 32; n1++)
            for (int c = 0; c < 2; c++) { const double v = hx[(size_t)f * SAMPLES + 32 * n1 + 2 * n2 + c]; nrm += v * v; }
        nrm = std::sqrt(nrm) + 1e-30;
        for (int o = 0; o < NOUT; o++) {
            double ref = 0.0, plain = 0.0;
            for (int k = 0; k < KDIM; k++) {
                const float a = hx[(size_t)f * SAMPLES + 32 * (k / 2) + 2 * n2 + (k & 1)];
                ref += (double)a * bd[(size_t)o * KDIM + k];
                uint32_t u; memcpy(&u, &a, 4); u &= 0xffffe000u; float ah; memcpy(&ah, &u, 4);
                plain += (double)ah * (double)bm[((size_t)(k / 32) * NOUT + o) * 32 + k % 32];      // what one TF32 pass would see
            }
            worst = std::max(worst, std::fabs((double)ho[(size_t)row * NOUT + o] - ref) / nrm);
            worst_plain = std::max(worst_plain, std::fabs(plain - ref) / nrm);
        }
    }
    printf("parity (1024 rows, incl. a frame with 80 dB dynamic range): split-TF32 max |err| / ||row|| = %.3e   (single TF32 pass would be %.3e)\n", worst, worst_plain);

    // ---- timing ----
    const char *names[5] = {"full pipeline", "MMA + readers (operands resident)", "MMA only", "loaders only", "readers only"};
    for (int mode = 0; mode < 5; mode++) {
        Params q{dx, dbm, dout, dsum, dcyc, tiles, mode, 0};
        CK(cudaMemset(dsum, 0, n_ctas * 8));
        dft_stage_kernel<<<n_ctas, THREADS, smem>>>(q);
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        dft_stage_kernel<<<n_ctas, THREADS, smem>>>(q);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        std::vector<long long> hc(n_ctas * 4);
        CK(cudaMemcpy(hc.data(), dcyc, hc.size() * 8, cudaMemcpyDeviceToHost));
        double c[4] = {0, 0, 0, 0};
        for (int b = 0; b < n_ctas; b++) for (int j = 0; j < 4; j++) c[j] += (double)hc[b * 4 + j] / n_ctas;
        const double fpc = (double)tiles * FRAMES_PER_TILE;
        printf("mode %d %-36s %.3f ms  %.1f M frames/s  cycles per frame per SM: total %.1f (loader %.1f, mma %.1f, reader %.1f)\n",
               mode, names[mode], ms, n_frames / (ms * 1e3), c[3] / fpc, c[0] / fpc, c[1] / fpc, c[2] / fpc);
    }
    printf("reference: feat_warp8 (FP32, all passes + mel + log + DCT) runs at ~395 cycles per frame per SM; its three FFT passes alone at ~270\n");
    return 0;
}
