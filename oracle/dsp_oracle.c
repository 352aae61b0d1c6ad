/*
 * dsp_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's src/dsp -> MFCC-retrieval hot path
 * (Audiofool934/dsp-final).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.  The
 * product path (dsp_final_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_pinned.py checks this file against
 *   (1) golden vectors produced by importing the reference itself
 *       (tests/golden/make_golden.py, run in the dev container),
 *   (2) the reference's own recorded known answers (librosa_compare.json STFT
 *       number, cache digests, table SHA-1s listed in SURVEY.md appendix A).
 *
 * Every function names the reference file:line it follows (paths relative to
 * the reference checkout).  Arithmetic is float64 exactly where the reference
 * is float64, float32 exactly where NumPy keeps float32 (pre-emphasis).
 * Build: see oracle/Makefile  (-O2 -ffp-contract=off: no FMA contraction, so
 * results do not depend on the host's FMA support).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define ORC_OK 0
#define ORC_EINVAL -1
#define ORC_ENOMEM -2

/* window ids shared with include/dspx.h */
#define ORC_WIN_HANN 0
#define ORC_WIN_HAMMING 1
#define ORC_WIN_RECT 2

/* ---- src/dsp/fft.py:7-10 (_next_pow_two) -------------------------------- */
int64_t orc_next_pow_two(int64_t n)
{
    int64_t p = 1;
    if (n <= 1) return 1;
    while (p < n) p <<= 1;
    return p;
}

/* ---- src/dsp/fft.py:13-24 (_bit_reverse_indices) ------------------------ */
static void bit_reverse_table(uint32_t *rev, int64_t n)
{
    int bits = 0;
    while (((int64_t)1 << bits) < n) bits++;
    for (int64_t i = 0; i < n; i++) {
        uint32_t b = (uint32_t)i, r = 0;
        for (int t = 0; t < bits; t++) {
            r = (r << 1) | (b & 1u);
            b >>= 1;
        }
        rev[i] = r;
    }
}

/*
 * ---- src/dsp/fft.py:27-61 (fft) ------------------------------------------
 * Forward DFT of a power-of-two length, radix-2 decimation in time, complex128.
 * re/im are overwritten.  Same schedule as the reference: permute by the
 * bit-reversal table, then for m = 2,4,..,n combine halves with
 * w_j = exp(-2*pi*i*j/m).  (np.exp of a purely imaginary complex128 is
 * (cos, sin) of the angle; libm cos/sin agree with NumPy's to the last ulp or
 * two, which is far below every tolerance used against this oracle.)
 */
int orc_fft_pow2(double *re, double *im, int64_t n)
{
    if (n < 1 || (n & (n - 1))) return ORC_EINVAL;
    if (n == 1) return ORC_OK;
    uint32_t *rev = (uint32_t *)malloc((size_t)n * sizeof(uint32_t));
    double *tr = (double *)malloc((size_t)n * sizeof(double));
    double *ti = (double *)malloc((size_t)n * sizeof(double));
    double *wr = (double *)malloc((size_t)(n / 2) * sizeof(double));
    double *wi = (double *)malloc((size_t)(n / 2) * sizeof(double));
    if (!rev || !tr || !ti || !wr || !wi) {
        free(rev); free(tr); free(ti); free(wr); free(wi);
        return ORC_ENOMEM;
    }
    bit_reverse_table(rev, n);          /* rebuilt per call, like the reference */
    for (int64_t i = 0; i < n; i++) { tr[i] = re[rev[i]]; ti[i] = im[rev[i]]; }
    memcpy(re, tr, (size_t)n * sizeof(double));
    memcpy(im, ti, (size_t)n * sizeof(double));

    for (int64_t m = 2; m <= n; m <<= 1) {
        int64_t half = m >> 1;
        double step = -2.0 * M_PI / (double)m;      /* angle = -2j*pi/m */
        for (int64_t j = 0; j < half; j++) {
            double a = (double)j * step;
            wr[j] = cos(a);
            wi[j] = sin(a);
        }
        for (int64_t k = 0; k < n; k += m) {
            for (int64_t j = 0; j < half; j++) {
                double xr = re[k + half + j], xi = im[k + half + j];
                double pr = wr[j] * xr - wi[j] * xi;  /* t = w * x[k+half+j] */
                double pi_ = wr[j] * xi + wi[j] * xr;
                double ur = re[k + j], ui = im[k + j];
                re[k + j] = ur + pr;            im[k + j] = ui + pi_;
                re[k + half + j] = ur - pr;     im[k + half + j] = ui - pi_;
            }
        }
    }
    free(rev); free(tr); free(ti); free(wr); free(wi);
    return ORC_OK;
}

/*
 * ---- src/dsp/fft.py:27-42 (length handling of fft) ------------------------
 * x has len_in complex samples.  n <= 0 means "n is None".  Truncate or
 * zero-pad to n, then zero-pad to the next power of two.  out_re/out_im must
 * hold orc_next_pow_two(n) values.  Returns the transform length (>0) or <0.
 */
int64_t orc_fft(const double *in_re, const double *in_im, int64_t len_in, int64_t n,
                double *out_re, double *out_im)
{
    if (n <= 0) n = len_in;
    if (n <= 0) return ORC_EINVAL;
    int64_t p = orc_next_pow_two(n);
    int64_t take = len_in < n ? len_in : n;
    for (int64_t i = 0; i < p; i++) {
        out_re[i] = i < take ? in_re[i] : 0.0;
        out_im[i] = (i < take && in_im) ? in_im[i] : 0.0;
    }
    int rc = orc_fft_pow2(out_re, out_im, p);
    return rc < 0 ? rc : p;
}

/* ---- src/dsp/fft.py:64-70 (ifft): conj(fft(conj x)) / N ------------------ */
int64_t orc_ifft(const double *in_re, const double *in_im, int64_t len_in, int64_t n,
                 double *out_re, double *out_im)
{
    if (n <= 0) n = len_in;
    if (n <= 0) return ORC_EINVAL;
    double *ci = (double *)malloc((size_t)len_in * sizeof(double));
    if (!ci) return ORC_ENOMEM;
    for (int64_t i = 0; i < len_in; i++) ci[i] = in_im ? -in_im[i] : 0.0;
    int64_t p = orc_fft(in_re, ci, len_in, n, out_re, out_im);
    free(ci);
    if (p < 0) return p;
    for (int64_t i = 0; i < p; i++) {
        out_re[i] = out_re[i] / (double)p;
        out_im[i] = -out_im[i] / (double)p;
    }
    return p;
}

/* ---- src/dsp/stft.py:12-24 (_get_window): periodic windows, float64 ------ */
int orc_window(int kind, int64_t frame_length, double *out)
{
    if (frame_length <= 0) return ORC_EINVAL;
    for (int64_t i = 0; i < frame_length; i++) {
        double c = cos(2.0 * M_PI * (double)i / (double)frame_length);
        if (kind == ORC_WIN_HANN) out[i] = 0.5 - 0.5 * c;
        else if (kind == ORC_WIN_HAMMING) out[i] = 0.54 - 0.46 * c;
        else if (kind == ORC_WIN_RECT) out[i] = 1.0;
        else return ORC_EINVAL;
    }
    return ORC_OK;
}

/* ---- src/dsp/stft.py:33 : n_frames = 1 + max(0,(L-fl)//hop) -------------- */
int64_t orc_num_frames(int64_t len, int64_t frame_length, int64_t hop_length)
{
    if (frame_length <= 0 || hop_length <= 0 || len < frame_length) return ORC_EINVAL;
    return 1 + (len - frame_length) / hop_length;
}

/*
 * ---- src/dsp/stft.py:43-56 (stft) + fft.py:73-77 (rfft) -------------------
 * signal: float64 [len].  Output: [n_frames, P/2+1] complex128 split into
 * out_re/out_im, P = next_pow2(n_fft) (fft.py:38-42).  Frame * window is
 * truncated to n_fft samples when n_fft < frame_length (fft.py:32-33), else
 * zero-padded.  No centring, no tail padding (stft.py:27-40).  Unlike the
 * reference (which reads past the buffer when len < frame_length, SURVEY
 * appendix A.5) this returns ORC_EINVAL for that case.
 */
int orc_stft(const double *signal, int64_t len, int64_t frame_length, int64_t hop_length,
             int window, int64_t n_fft, double *out_re, double *out_im)
{
    int64_t T = orc_num_frames(len, frame_length, hop_length);
    if (T < 0) return ORC_EINVAL;
    if (n_fft <= 0) n_fft = frame_length;
    int64_t P = orc_next_pow_two(n_fft);
    int64_t bins = P / 2 + 1;
    int64_t take = frame_length < n_fft ? frame_length : n_fft;
    double *win = (double *)malloc((size_t)frame_length * sizeof(double));
    double *br = (double *)malloc((size_t)P * sizeof(double));
    double *bi = (double *)malloc((size_t)P * sizeof(double));
    if (!win || !br || !bi) { free(win); free(br); free(bi); return ORC_ENOMEM; }
    int rc = orc_window(window, frame_length, win);
    for (int64_t t = 0; rc == ORC_OK && t < T; t++) {
        const double *f = signal + t * hop_length;
        for (int64_t i = 0; i < P; i++) {
            br[i] = i < take ? f[i] * win[i] : 0.0;
            bi[i] = 0.0;
        }
        rc = orc_fft_pow2(br, bi, P);
        for (int64_t k = 0; k < bins; k++) {
            out_re[t * bins + k] = br[k];
            out_im[t * bins + k] = bi[k];
        }
    }
    free(win); free(br); free(bi);
    return rc;
}

/* ---- src/dsp/mfcc.py:24-29 (HTK mel scale) ------------------------------- */
static double hz_to_mel(double hz) { return 2595.0 * log10(1.0 + hz / 700.0); }
static double mel_to_hz(double mel) { return 700.0 * (pow(10.0, mel / 2595.0) - 1.0); }

/*
 * ---- src/dsp/mfcc.py:32-58 (mel_filterbank) --------------------------------
 * out: float64 [n_mels, n_fft/2+1], zero-filled here.  f_max < 0 means None.
 * np.linspace(a, b, n) = a + i*((b-a)/(n-1)) with the last point forced to b.
 */
int orc_mel_filterbank(int n_mels, int64_t n_fft, int sample_rate, double f_min, double f_max,
                       double *out)
{
    if (n_mels <= 0 || n_fft <= 0 || sample_rate <= 0) return ORC_EINVAL;
    if (f_max < 0.0) f_max = (double)sample_rate / 2.0;
    int64_t bins = n_fft / 2 + 1;
    int np_ = n_mels + 2;
    int64_t *bf = (int64_t *)malloc((size_t)np_ * sizeof(int64_t));
    if (!bf) return ORC_ENOMEM;
    double mel_lo = hz_to_mel(f_min), mel_hi = hz_to_mel(f_max);
    double step = (mel_hi - mel_lo) / (double)(np_ - 1);
    for (int i = 0; i < np_; i++) {
        double mel = (i == np_ - 1) ? mel_hi : (double)i * step + mel_lo;
        double hz = mel_to_hz(mel);
        bf[i] = (int64_t)floor((double)(n_fft + 1) * hz / (double)sample_rate);
    }
    memset(out, 0, (size_t)n_mels * (size_t)bins * sizeof(double));
    for (int m = 1; m <= n_mels; m++) {
        int64_t left = bf[m - 1], center = bf[m], right = bf[m + 1];
        if (right <= left) continue;
        int64_t up = center - left > 1 ? center - left : 1;
        int64_t dn = right - center > 1 ? right - center : 1;
        for (int64_t k = left; k < center; k++)
            if (k >= 0 && k < bins) out[(m - 1) * bins + k] = (double)(k - left) / (double)up;
        for (int64_t k = center; k < right; k++)
            if (k >= 0 && k < bins) out[(m - 1) * bins + k] = (double)(right - k) / (double)dn;
    }
    free(bf);
    return ORC_OK;
}

/* ---- src/dsp/mfcc.py:79-83 (_dct_basis): C[k,n] = cos(pi/N*(n+0.5)*k) ----- */
int orc_dct_basis(int n_mfcc, int n, double *out)
{
    if (n_mfcc <= 0 || n <= 0) return ORC_EINVAL;
    for (int k = 0; k < n_mfcc; k++)
        for (int j = 0; j < n; j++)
            out[k * n + j] = cos(M_PI / (double)n * ((double)j + 0.5) * (double)k);
    return ORC_OK;
}

typedef struct {
    int sample_rate;
    int64_t frame_length;
    int64_t hop_length;
    int64_t n_fft;          /* <= 0: None -> frame_length (mfcc.py:89) */
    int n_mels;
    int n_mfcc;
    double f_min;
    double f_max;           /* < 0: None -> sample_rate/2 */
    double pre_emphasis;
    int window;
} orc_config;

/*
 * ---- src/dsp/mfcc.py:86-103 (log_mel_spectrogram) --------------------------
 * signal_f32 != NULL: float32 input, pre-emphasis evaluated in float32 (NumPy
 * keeps float32 for  signal[1:] - alpha*signal[:-1]  with a Python-float alpha:
 * alpha is rounded to float32, the product and the difference are each rounded
 * to float32).  Otherwise signal_f64 is used and everything is float64.
 * Output: float64 [n_frames, n_mels].
 */
int orc_log_mel(const float *signal_f32, const double *signal_f64, int64_t len,
                const orc_config *cfg, double *out)
{
    int64_t T = orc_num_frames(len, cfg->frame_length, cfg->hop_length);
    if (T < 0) return ORC_EINVAL;
    int64_t n_fft = cfg->n_fft > 0 ? cfg->n_fft : cfg->frame_length;
    n_fft = orc_next_pow_two(n_fft);            /* mfcc.py:91 */
    int64_t bins = n_fft / 2 + 1;
    double *y = (double *)malloc((size_t)len * sizeof(double));
    double *sr = (double *)malloc((size_t)T * (size_t)bins * sizeof(double));
    double *si = (double *)malloc((size_t)T * (size_t)bins * sizeof(double));
    double *fb = (double *)malloc((size_t)cfg->n_mels * (size_t)bins * sizeof(double));
    int rc = ORC_OK;
    if (!y || !sr || !si || !fb) { rc = ORC_ENOMEM; goto done; }

    if (signal_f32) {
        if (cfg->pre_emphasis > 0.0) {          /* mfcc.py:87-88, float32 arithmetic */
            volatile float a = (float)cfg->pre_emphasis;
            y[0] = (double)signal_f32[0];
            for (int64_t i = 1; i < len; i++) {
                volatile float prod = a * signal_f32[i - 1];
                volatile float diff = signal_f32[i] - prod;
                y[i] = (double)diff;
            }
        } else {
            for (int64_t i = 0; i < len; i++) y[i] = (double)signal_f32[i];
        }
    } else {
        if (cfg->pre_emphasis > 0.0) {
            y[0] = signal_f64[0];
            for (int64_t i = 1; i < len; i++)
                y[i] = signal_f64[i] - cfg->pre_emphasis * signal_f64[i - 1];
        } else {
            memcpy(y, signal_f64, (size_t)len * sizeof(double));
        }
    }
    rc = orc_stft(y, len, cfg->frame_length, cfg->hop_length, cfg->window, n_fft, sr, si);
    if (rc != ORC_OK) goto done;
    rc = orc_mel_filterbank(cfg->n_mels, n_fft, cfg->sample_rate, cfg->f_min, cfg->f_max, fb);
    if (rc != ORC_OK) goto done;
    for (int64_t t = 0; t < T; t++) {
        for (int m = 0; m < cfg->n_mels; m++) {
            double acc = 0.0;                   /* power @ fbank.T, mfcc.py:99-101 */
            const double *w = fb + (int64_t)m * bins;
            for (int64_t k = 0; k < bins; k++) {
                if (w[k] == 0.0) continue;      /* exact: adding 0*p changes nothing */
                double a = sr[t * bins + k], b = si[t * bins + k];
                double mag = hypot(a, b);       /* np.abs(spec) ** 2 */
                acc += (mag * mag) * w[k];
            }
            if (acc < 1e-10) acc = 1e-10;       /* mfcc.py:102 */
            out[t * cfg->n_mels + m] = log(acc);/* mfcc.py:103 */
        }
    }
done:
    free(y); free(sr); free(si); free(fb);
    return rc;
}

/* ---- src/dsp/mfcc.py:73-76,106-109 (dct_type_2, mfcc) --------------------- */
int orc_mfcc(const float *signal_f32, const double *signal_f64, int64_t len,
             const orc_config *cfg, double *out_mfcc, double *out_logmel_or_null)
{
    int64_t T = orc_num_frames(len, cfg->frame_length, cfg->hop_length);
    if (T < 0 || cfg->n_mfcc <= 0) return ORC_EINVAL;
    double *lm = out_logmel_or_null;
    if (!lm) lm = (double *)malloc((size_t)T * (size_t)cfg->n_mels * sizeof(double));
    double *basis = (double *)malloc((size_t)cfg->n_mfcc * (size_t)cfg->n_mels * sizeof(double));
    int rc = (!lm || !basis) ? ORC_ENOMEM : orc_log_mel(signal_f32, signal_f64, len, cfg, lm);
    if (rc == ORC_OK) rc = orc_dct_basis(cfg->n_mfcc, cfg->n_mels, basis);
    if (rc == ORC_OK) {
        for (int64_t t = 0; t < T; t++)
            for (int k = 0; k < cfg->n_mfcc; k++) {
                double acc = 0.0;
                for (int j = 0; j < cfg->n_mels; j++)
                    acc += lm[t * cfg->n_mels + j] * basis[k * cfg->n_mels + j];
                out_mfcc[t * cfg->n_mfcc + k] = 2.0 * acc;
            }
    }
    if (!out_logmel_or_null) free(lm);
    free(basis);
    return rc;
}

/*
 * ---- src/retrieval/retrieval.py:19-23 and :38-41 (clip embedding) ----------
 * concat(mean over frames, population std over frames), float64.
 */
int orc_embedding(const double *feats, int64_t T, int C, double *out)
{
    if (T <= 0 || C <= 0) return ORC_EINVAL;
    for (int c = 0; c < C; c++) {
        double s = 0.0;
        for (int64_t t = 0; t < T; t++) s += feats[t * C + c];
        double mean = s / (double)T;
        double v = 0.0;
        for (int64_t t = 0; t < T; t++) {
            double d = feats[t * C + c] - mean;
            v += d * d;
        }
        out[c] = mean;
        out[C + c] = sqrt(v / (double)T);
    }
    return ORC_OK;
}

/*
 * ---- CMVN option (no reference counterpart: src/dsp/mfcc.py:102-109 end at log and DCT) ----
 * BASELINE.json's north_star names a log/CMVN epilogue; this is the CPU statement of the option the
 * GPU side offers (dspx_cmvn): y = (x - mean_t) / (std_t + eps), population std, float64, in place.
 */
int orc_cmvn(double *feats, int64_t T, int C, double eps)
{
    if (T <= 0 || C <= 0) return ORC_EINVAL;
    for (int c = 0; c < C; c++) {
        double s = 0.0;
        for (int64_t t = 0; t < T; t++) s += feats[t * C + c];
        const double mean = s / (double)T;
        double v = 0.0;
        for (int64_t t = 0; t < T; t++) {
            const double d = feats[t * C + c] - mean;
            v += d * d;
        }
        const double r = 1.0 / (sqrt(v / (double)T) + eps);
        for (int64_t t = 0; t < T; t++) feats[t * C + c] = (feats[t * C + c] - mean) * r;
    }
    return ORC_OK;
}

/*
 * Batched front end used for the CPU baseline and for large parity sets:
 * clips float32 [B, len] -> mfcc float32 [B,T,n_mfcc] (the cache dtype,
 * src/features/cache.py:74) and/or log-mel float32 [B,T,n_mels] and/or
 * embeddings float32 [B, 2*n_mfcc] computed from the float32-rounded MFCCs
 * (the cached retrieval path, retrieval.py:38-41).  One clip per task, all
 * host threads -- the same fan-out as scripts/tools/precompute_features.py:96-107.
 */
int orc_features_batch(const float *clips, int64_t n_clips, int64_t len, int64_t stride,
                       const orc_config *cfg, float *mfcc_out, float *logmel_out,
                       float *embed_out, int n_threads)
{
    int64_t T = orc_num_frames(len, cfg->frame_length, cfg->hop_length);
    if (T < 0) return ORC_EINVAL;
    int rc_all = ORC_OK;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int64_t b = 0; b < n_clips; b++) {
        double *mf = (double *)malloc((size_t)T * (size_t)cfg->n_mfcc * sizeof(double));
        double *lm = (double *)malloc((size_t)T * (size_t)cfg->n_mels * sizeof(double));
        int rc = (!mf || !lm) ? ORC_ENOMEM : orc_mfcc(clips + b * stride, NULL, len, cfg, mf, lm);
        if (rc == ORC_OK) {
            int64_t nm = T * cfg->n_mfcc, nl = T * cfg->n_mels;
            if (mfcc_out) for (int64_t i = 0; i < nm; i++) mfcc_out[b * nm + i] = (float)mf[i];
            if (logmel_out) for (int64_t i = 0; i < nl; i++) logmel_out[b * nl + i] = (float)lm[i];
            if (embed_out) {
                double e[2 * 1024];
                if (cfg->n_mfcc > 1024) rc = ORC_EINVAL;
                else {
                    for (int64_t i = 0; i < nm; i++) mf[i] = (double)(float)mf[i];
                    orc_embedding(mf, T, cfg->n_mfcc, e);
                    for (int c = 0; c < 2 * cfg->n_mfcc; c++)
                        embed_out[b * 2 * cfg->n_mfcc + c] = (float)e[c];
                }
            }
        }
        if (rc != ORC_OK) {
#ifdef _OPENMP
#pragma omp critical
#endif
            rc_all = rc;
        }
        free(mf); free(lm);
    }
    return rc_all;
}

/*
 * ---- src/retrieval/retrieval.py:46-49 (cosine_similarity) + :65 (top-k) ----
 * q [nq,d], db [ndb,d] float64.  Rows are scaled by 1/(||row|| + 1e-10) (eps
 * added to the norm, not under the root), scores are left-to-right chains of
 * IEEE fused multiply-adds in float64 (fma(): the same operation sequence the
 * CUDA kernel issues, so scores match it bit for bit), and the k best per query are
 * returned in descending score order with ties broken by the LOWER database
 * index -- i.e. np.argsort(-sims, axis=1, kind="stable")[:, :k].  (The
 * reference's default introsort leaves tie order unspecified)
 */
int orc_cosine_topk(const double *q, int64_t nq, const double *db, int64_t ndb, int d, int k,
                    int32_t *idx_out, double *score_out_or_null, int n_threads)
{
    if (nq < 0 || ndb <= 0 || d <= 0 || k <= 0 || k > ndb) return ORC_EINVAL;
    double *dbn = (double *)malloc((size_t)ndb * (size_t)d * sizeof(double));
    if (!dbn) return ORC_ENOMEM;
    for (int64_t j = 0; j < ndb; j++) {
        double s = 0.0;
        for (int c = 0; c < d; c++) s = fma(db[j * d + c], db[j * d + c], s);
        double inv = sqrt(s) + 1e-10;
        for (int c = 0; c < d; c++) dbn[j * d + c] = db[j * d + c] / inv;
    }
    int rc_all = ORC_OK;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 8)
#endif
    for (int64_t i = 0; i < nq; i++) {
        double *qn = (double *)malloc((size_t)d * sizeof(double));
        double *bs = (double *)malloc((size_t)k * sizeof(double));
        int32_t *bi = (int32_t *)malloc((size_t)k * sizeof(int32_t));
        if (!qn || !bs || !bi) { rc_all = ORC_ENOMEM; free(qn); free(bs); free(bi); continue; }
        double s = 0.0;
        for (int c = 0; c < d; c++) s = fma(q[i * d + c], q[i * d + c], s);
        double inv = sqrt(s) + 1e-10;
        for (int c = 0; c < d; c++) qn[c] = q[i * d + c] / inv;
        int have = 0;
        for (int64_t j = 0; j < ndb; j++) {
            double acc = 0.0;
            for (int c = 0; c < d; c++) acc = fma(qn[c], dbn[j * d + c], acc);
            /* sorted insert; strict > keeps the earlier (lower) index first on ties */
            if (have < k || acc > bs[have - 1]) {
                int pos = have < k ? have : k - 1;
                while (pos > 0 && acc > bs[pos - 1]) {
                    bs[pos] = bs[pos - 1];
                    bi[pos] = bi[pos - 1];
                    pos--;
                }
                bs[pos] = acc;
                bi[pos] = (int32_t)j;
                if (have < k) have++;
            }
        }
        for (int r = 0; r < k; r++) {
            idx_out[i * k + r] = bi[r];
            if (score_out_or_null) score_out_or_null[i * k + r] = bs[r];
        }
        free(qn); free(bs); free(bi);
    }
    free(dbn);
    return rc_all;
}

/* ---- src/retrieval/retrieval.py:66-70 : hit@k counted over queries -------- */
int64_t orc_hits_at_k(const int32_t *topk_idx, int64_t nq, int k_stride, int k,
                      const int32_t *targets_db, const int32_t *targets_q)
{
    int64_t hits = 0;
    for (int64_t i = 0; i < nq; i++) {
        int hit = 0;
        for (int r = 0; r < k && !hit; r++)
            hit = targets_db[topk_idx[i * k_stride + r]] == targets_q[i];
        hits += hit;
    }
    return hits;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
