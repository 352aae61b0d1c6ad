// feat_generic.cuh -- generic fused feature kernel: any power-of-two transform
// length P in [16, 8192], any frame/hop, complex STFT or log-mel/MFCC output.
//
// Reference ops fused here (one launch, nothing but the outputs touches HBM):
//   pre-emphasis           src/dsp/mfcc.py:87-88     (float32, rounded mul then sub)
//   framing                src/dsp/stft.py:27-40     (no centring, no tail pad)
//   window                 src/dsp/stft.py:12-24,56
//   zero-pad / truncate    src/dsp/fft.py:32-42
//   FFT                    src/dsp/fft.py:27-61,73-77 -> Stockham radix-4/2 in shared memory on
//                          the N/2-point packed-real transform (+ split post-pass)
//   |X|^2                  src/dsp/mfcc.py:99
//   mel projection         src/dsp/mfcc.py:100-101   (sparse rows of the same table)
//   log(max(.,1e-10))      src/dsp/mfcc.py:102-103
//   DCT-II x2              src/dsp/mfcc.py:73-83,108
//
// A CTA owns a run of consecutive frames of one clip and pushes G frames at a
// time through shared memory.  The work between two barriers is written as
// "phase" functions over a flat item index so that csrc/emu.cu can replay the
// very same code on the CPU (tests only; the product never runs it there).
#pragma once

#include "dspx_internal.cuh"

namespace dspx {

constexpr int GEN_THREADS = 256;
constexpr int GEN_MEL_PARTS = 8;     // partial sums per mel filter (fixed order -> deterministic)

struct GenParams {
    const float *clips;
    int64_t n_clips, clip_len, clip_stride, n_frames;
    int frame_length, hop, take, P, M, n_bins;
    int n_stages;
    int radix[16];
    int n_mels, n_mfcc;
    int G;                  // frames in flight per CTA
    int frames_per_cta;     // multiple of G
    int ctas_per_clip;
    int pre;                // apply pre-emphasis
    float alpha;
    const float *window;
    const float2 *tw;       // exp(-2 pi i j / P), j in [0, P)
    const int32_t *fb_start, *fb_cnt, *fb_off;
    const float *fb_w;
    const float *dct2;
    float *logmel;          // may be null
    int64_t lm_ts, lm_fs;   // log-mel strides between frames / between filters ([B,T,M]: M,1; [B,1,M,T]: 1,T)
    float *mfcc;            // may be null
    float2 *stft;           // non-null selects complex-spectrum output (no mel stages)
};

// shared-memory carve-up (floats): two ping-pong complex buffers, partial mel sums, log-mel row
struct GenSmem {
    float2 *a, *b;
    float *part, *lm;
};

DSPX_HD size_t gen_smem_bytes(int G, int M, int n_mels)
{
    return (size_t)G * M * sizeof(float2) * 2 + (size_t)G * n_mels * (GEN_MEL_PARTS + 1) * sizeof(float);
}

DSPX_HD GenSmem gen_carve(void *base, int G, int M, int n_mels)
{
    GenSmem s;
    s.a = reinterpret_cast<float2 *>(base);
    s.b = s.a + (size_t)G * M;
    s.part = reinterpret_cast<float *>(s.b + (size_t)G * M);
    s.lm = s.part + (size_t)G * n_mels * GEN_MEL_PARTS;
    return s;
}

// pre-emphasised sample y[s] of one clip (global sample index s)
DSPX_HD float gen_sample(const GenParams &p, const float *clip, int64_t s)
{
    float x = clip[s];
    if (p.pre && s > 0) x = DSPX_FSUB_RN(x, DSPX_FMUL_RN(p.alpha, clip[s - 1]));
    return x;
}

// phase 1: items [0, G*M): pack windowed samples (2m, 2m+1) of frame g into a[g][m]
DSPX_HD void gen_phase_load(const GenParams &p, const GenSmem &sm, const float *clip, int64_t t0, int i)
{
    const int g = i / p.M, m = i - g * p.M;
    const int64_t t = t0 + g;
    float2 v = make_float2(0.f, 0.f);
    if (t < p.n_frames) {
        const int n0 = 2 * m;
        const int64_t s0 = t * (int64_t)p.hop + n0;
        if (n0 < p.take) v.x = gen_sample(p, clip, s0) * p.window[n0];
        if (n0 + 1 < p.take) v.y = gen_sample(p, clip, s0 + 1) * p.window[n0 + 1];
    }
    sm.a[(size_t)g * p.M + m] = v;
}

DSPX_HD float2 cmul(float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }

// phase 2.s: items [0, G*M/R): one radix-R Stockham butterfly (autosort, natural order out)
DSPX_HD void gen_phase_stage(const GenParams &p, const float2 *in, float2 *out, int stage, int ns, int i)
{
    const int R = p.radix[stage];
    const int per = p.M / R;
    const int g = i / per, j = i - g * per;
    const int k = j & (ns - 1);
    const float2 *src = in + (size_t)g * p.M;
    float2 *dst = out + (size_t)g * p.M;
    const int tstep = (p.P / (ns * R)) * k;          // table index of exp(-2 pi i k / (ns R))
    const int j0 = ((j - k) * R) + k;                // (j / ns) * ns * R + k
    if (R == 4) {
        float2 v0 = src[j], v1 = src[j + per], v2 = src[j + 2 * per], v3 = src[j + 3 * per];
        if (ns > 1) {
            v1 = cmul(v1, p.tw[tstep]);
            v2 = cmul(v2, p.tw[2 * tstep]);
            v3 = cmul(v3, p.tw[3 * tstep]);
        }
        const float2 s02 = make_float2(v0.x + v2.x, v0.y + v2.y), d02 = make_float2(v0.x - v2.x, v0.y - v2.y);
        const float2 s13 = make_float2(v1.x + v3.x, v1.y + v3.y);
        const float2 d13 = make_float2(v1.y - v3.y, -(v1.x - v3.x));      // (v1 - v3) * (-i)
        dst[j0] = make_float2(s02.x + s13.x, s02.y + s13.y);
        dst[j0 + ns] = make_float2(d02.x + d13.x, d02.y + d13.y);
        dst[j0 + 2 * ns] = make_float2(s02.x - s13.x, s02.y - s13.y);
        dst[j0 + 3 * ns] = make_float2(d02.x - d13.x, d02.y - d13.y);
    } else {
        float2 v0 = src[j], v1 = src[j + per];
        if (ns > 1) v1 = cmul(v1, p.tw[tstep]);
        dst[j0] = make_float2(v0.x + v1.x, v0.y + v1.y);
        dst[j0 + ns] = make_float2(v0.x - v1.x, v0.y - v1.y);
    }
}

// phase 3: items [0, G*(M/2+1)): split the packed transform into bins k and M-k
//   E = (Z[k] + conj Z[M-k]) / 2,  O = (Z[k] - conj Z[M-k]) / (2i),
//   X[k] = E + W^k O,  X[M-k] = conj(E - W^k O),  W = exp(-2 pi i / P)
DSPX_HD void gen_phase_post(const GenParams &p, const float2 *fin, float *pw, int64_t clip_idx, int64_t t0, int i)
{
    const int cnt = p.M / 2 + 1;
    const int g = i / cnt, k = i - g * cnt;
    const int64_t t = t0 + g;
    if (t >= p.n_frames) return;
    const float2 zk = fin[(size_t)g * p.M + k];
    const float2 zm = fin[(size_t)g * p.M + ((p.M - k) & (p.M - 1))];
    const float er = 0.5f * (zk.x + zm.x), ei = 0.5f * (zk.y - zm.y);
    const float orr = 0.5f * (zk.y + zm.y), oi = -0.5f * (zk.x - zm.x);
    const float2 w = p.tw[k];
    const float tr = w.x * orr - w.y * oi, ti = w.x * oi + w.y * orr;
    const float2 xa = make_float2(er + tr, ei + ti);          // X[k]
    const float2 xb = make_float2(er - tr, -(ei - ti));       // X[M-k]
    if (p.stft) {
        float2 *row = p.stft + ((size_t)clip_idx * p.n_frames + t) * p.n_bins;
        row[k] = xa;
        if (p.M - k != k) row[p.M - k] = xb;
    } else {
        float *row = pw + (size_t)g * (p.M + 1);
        row[k] = xa.x * xa.x + xa.y * xa.y;
        if (p.M - k != k) row[p.M - k] = xb.x * xb.x + xb.y * xb.y;
    }
}

// phase 4: items [0, G*n_mels*PARTS): strided partial sums of one sparse filterbank row
DSPX_HD void gen_phase_melpart(const GenParams &p, const float *pw, float *part, int i)
{
    const int per = p.n_mels * GEN_MEL_PARTS;
    const int g = i / per, r = i - g * per;
    const int f = r / GEN_MEL_PARTS, q0 = r - f * GEN_MEL_PARTS;
    const float *row = pw + (size_t)g * (p.M + 1) + p.fb_start[f];
    const float *w = p.fb_w + p.fb_off[f];
    const int cnt = p.fb_cnt[f];
    float acc = 0.f;
    for (int q = q0; q < cnt; q += GEN_MEL_PARTS) acc = fmaf(w[q], row[q], acc);
    part[i] = acc;
}

// phase 5: items [0, G*n_mels): finish the row sum, floor, log, store
DSPX_HD void gen_phase_logmel(const GenParams &p, const float *part, float *lm, int64_t clip_idx, int64_t t0, int i)
{
    const int g = i / p.n_mels, f = i - g * p.n_mels;
    const float *q = part + (size_t)i * GEN_MEL_PARTS;
    float s = 0.f;
    for (int r = 0; r < GEN_MEL_PARTS; r++) s += q[r];
    const float v = logf(fmaxf(s, 1e-10f));
    lm[i] = v;
    const int64_t t = t0 + g;
    if (p.logmel && t < p.n_frames) p.logmel[(size_t)clip_idx * p.n_frames * p.n_mels + t * p.lm_ts + f * p.lm_fs] = v;
}

// phase 6: items [0, G*n_mfcc): DCT-II (table already carries the factor 2)
DSPX_HD void gen_phase_dct(const GenParams &p, const float *lm, int64_t clip_idx, int64_t t0, int i)
{
    const int g = i / p.n_mfcc, c = i - g * p.n_mfcc;
    const int64_t t = t0 + g;
    if (t >= p.n_frames) return;
    const float *row = lm + (size_t)g * p.n_mels;
    const float *basis = p.dct2 + (size_t)c * p.n_mels;
    float acc = 0.f;
    for (int f = 0; f < p.n_mels; f++) acc = fmaf(row[f], basis[f], acc);
    p.mfcc[((size_t)clip_idx * p.n_frames + t) * p.n_mfcc + c] = acc;
}

#if defined(__CUDACC__)
__global__ void __launch_bounds__(GEN_THREADS) feat_generic_kernel(const GenParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GenSmem sm = gen_carve(smem_raw, p.G, p.M, p.n_mels);
    const int64_t clip_idx = blockIdx.x / p.ctas_per_clip;
    const int64_t chunk = blockIdx.x - clip_idx * p.ctas_per_clip;
    const float *clip = p.clips + clip_idx * p.clip_stride;
    const int64_t t_begin = chunk * p.frames_per_cta;
    int64_t t_end = t_begin + p.frames_per_cta;
    if (t_end > p.n_frames) t_end = p.n_frames;
    const int tid = threadIdx.x;

    for (int64_t t0 = t_begin; t0 < t_end; t0 += p.G) {
        for (int i = tid; i < p.G * p.M; i += GEN_THREADS) gen_phase_load(p, sm, clip, t0, i);
        __syncthreads();
        float2 *in = sm.a, *out = sm.b;
        int ns = 1;
        for (int s = 0; s < p.n_stages; s++) {
            const int items = p.G * (p.M / p.radix[s]);
            for (int i = tid; i < items; i += GEN_THREADS) gen_phase_stage(p, in, out, s, ns, i);
            __syncthreads();
            ns *= p.radix[s];
            float2 *tmp = in; in = out; out = tmp;
        }
        // transform now in `in`; `out` is free and doubles as the power-spectrum buffer
        float *pw = reinterpret_cast<float *>(out);
        for (int i = tid; i < p.G * (p.M / 2 + 1); i += GEN_THREADS) gen_phase_post(p, in, pw, clip_idx, t0, i);
        __syncthreads();
        if (!p.stft) {
            for (int i = tid; i < p.G * p.n_mels * GEN_MEL_PARTS; i += GEN_THREADS) gen_phase_melpart(p, pw, sm.part, i);
            __syncthreads();
            for (int i = tid; i < p.G * p.n_mels; i += GEN_THREADS) gen_phase_logmel(p, sm.part, sm.lm, clip_idx, t0, i);
            __syncthreads();
            if (p.mfcc)
                for (int i = tid; i < p.G * p.n_mfcc; i += GEN_THREADS) gen_phase_dct(p, sm.lm, clip_idx, t0, i);
        }
        // the next load phase writes sm.a, which post (if a was `in`) finished reading
        // before the barrier above; part/lm are rewritten only after two more barriers.
        __syncthreads();
    }
}
#endif

}  // namespace dspx
