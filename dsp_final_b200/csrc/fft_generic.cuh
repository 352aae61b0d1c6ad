// fft_generic.cuh -- batched complex FFT of any power-of-two length in global
// memory, for the fft / ifft / rfft entry points (src/dsp/fft.py:27-77).
//
// The reference is an iterative radix-2 DIT with a bit-reversal pass; here the
// same transform is a Stockham autosort (no permutation pass), radix 4 with one
// radix-2 stage when log2(P) is odd, ping-ponging between out/work so that the
// last stage lands in `out`.  Stage 0 reads the caller's buffer and applies the
// truncate / zero-pad rule of fft.py:32-42.  Twiddles come from sincospi in
// float64 (this path is API completeness, not the throughput path; the fused
// feature kernels carry their own tables).
#pragma once

#include "dspx_internal.cuh"

namespace dspx {

struct FftStage {
    const float2 *in;
    float2 *out;
    int64_t batch;
    int64_t in_stride;   // row stride of `in` in complex elements
    int64_t n_valid;     // stage 0: samples taken from each input row (rest read as zero)
    int64_t P;
    int64_t ns;          // product of earlier radices
    int R;               // 4 or 2
    int inverse;
    int first, last;
};

#if defined(__CUDACC__)
__device__ __forceinline__ float2 fft_fetch(const FftStage &s, const float2 *row, int64_t idx)
{
    if (s.first && idx >= s.n_valid) return make_float2(0.f, 0.f);
    return row[idx];
}

__global__ void __launch_bounds__(256) fft_stage_kernel(const FftStage s)
{
    const int64_t per = s.P / s.R;
    const int64_t total = s.batch * per;
    for (int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gi < total;
         gi += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = gi / per, j = gi - b * per;
        const int64_t k = j & (s.ns - 1);
        const float2 *row = s.in + b * s.in_stride;
        float2 *dst = s.out + b * s.P;
        const int64_t j0 = (j - k) * s.R + k;
        double sn, cs;
        sincospi(-2.0 * (double)k / (double)(s.ns * s.R), &sn, &cs);
        if (s.inverse) sn = -sn;
        const float scale = (s.last && s.inverse) ? (float)(1.0 / (double)s.P) : 1.0f;
        if (s.R == 4) {
            float2 v0 = fft_fetch(s, row, j), v1 = fft_fetch(s, row, j + per);
            float2 v2 = fft_fetch(s, row, j + 2 * per), v3 = fft_fetch(s, row, j + 3 * per);
            if (s.ns > 1) {
                const double c2 = cs * cs - sn * sn, s2 = 2.0 * cs * sn;
                const double c3 = c2 * cs - s2 * sn, s3 = c2 * sn + s2 * cs;
                v1 = cmul(v1, make_float2((float)cs, (float)sn));
                v2 = cmul(v2, make_float2((float)c2, (float)s2));
                v3 = cmul(v3, make_float2((float)c3, (float)s3));
            }
            const float2 a = make_float2(v0.x + v2.x, v0.y + v2.y), bq = make_float2(v0.x - v2.x, v0.y - v2.y);
            const float2 c = make_float2(v1.x + v3.x, v1.y + v3.y);
            float2 d = make_float2(v1.y - v3.y, -(v1.x - v3.x));         // (v1 - v3) * (-i)
            if (s.inverse) d = make_float2(-d.x, -d.y);                  // (v1 - v3) * (+i)
            dst[j0] = make_float2(scale * (a.x + c.x), scale * (a.y + c.y));
            dst[j0 + s.ns] = make_float2(scale * (bq.x + d.x), scale * (bq.y + d.y));
            dst[j0 + 2 * s.ns] = make_float2(scale * (a.x - c.x), scale * (a.y - c.y));
            dst[j0 + 3 * s.ns] = make_float2(scale * (bq.x - d.x), scale * (bq.y - d.y));
        } else {
            float2 v0 = fft_fetch(s, row, j), v1 = fft_fetch(s, row, j + per);
            if (s.ns > 1) v1 = cmul(v1, make_float2((float)cs, (float)sn));
            dst[j0] = make_float2(scale * (v0.x + v1.x), scale * (v0.y + v1.y));
            dst[j0 + s.ns] = make_float2(scale * (v0.x - v1.x), scale * (v0.y - v1.y));
        }
    }
}

// P == 1: fft of one sample is the sample (fft.py:44-45)
__global__ void fft_copy1_kernel(const float2 *in, float2 *out, int64_t batch, int64_t in_stride, int64_t n_valid)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch) out[b] = n_valid > 0 ? in[b * in_stride] : make_float2(0.f, 0.f);
}
#endif

}  // namespace dspx
