// tables.cuh -- host-side constant tables of a plan, float64 like the reference.
//
//   window      src/dsp/stft.py:12-24   periodic hann / hamming / rect
//   mel bank    src/dsp/mfcc.py:24-58   HTK mel scale, floor-binned, unnormalised triangles
//   DCT basis   src/dsp/mfcc.py:73-83   cos(pi/N (n+0.5) k), output scaled by 2
//
// The dense filterbank is what the reference multiplies by (power @ fbank.T);
// the kernels use two sparse views of exactly the same numbers:
//   row view   per filter: [start, start+cnt) and its weights      (generic kernel)
//   bin view   per bin: the (at most two, adjacent) filters it feeds (warp8 kernel)
#pragma once

#include "dspx_internal.cuh"

namespace dspx {

inline int64_t next_pow_two(int64_t n)
{
    int64_t p = 1;
    while (p < n) p <<= 1;
    return p;
}

inline double hz_to_mel(double hz) { return 2595.0 * std::log10(1.0 + hz / 700.0); }
inline double mel_to_hz(double mel) { return 700.0 * (std::pow(10.0, mel / 2595.0) - 1.0); }

inline bool build_window(int kind, int n, std::vector<double> &w)
{
    w.resize(n);
    for (int i = 0; i < n; i++) {
        double c = std::cos(2.0 * M_PI * (double)i / (double)n);
        switch (kind) {
            case DSPX_WINDOW_HANN: w[i] = 0.5 - 0.5 * c; break;
            case DSPX_WINDOW_HAMMING: w[i] = 0.54 - 0.46 * c; break;
            case DSPX_WINDOW_RECT: w[i] = 1.0; break;
            default: return false;
        }
    }
    return true;
}

// n_mels + 2 points evenly spaced in mel (np.linspace: start + i*step, last = stop),
// mapped back to Hz and floored onto FFT bins with the reference's (n_fft + 1) factor.
inline void mel_bin_edges(int n_mels, int n_fft, int sr, double f_min, double f_max,
                          std::vector<int64_t> &edges)
{
    const int np = n_mels + 2;
    edges.resize(np);
    const double lo = hz_to_mel(f_min), hi = hz_to_mel(f_max);
    const double step = (hi - lo) / (double)(np - 1);
    for (int i = 0; i < np; i++) {
        double mel = (i == np - 1) ? hi : (double)i * step + lo;
        edges[i] = (int64_t)std::floor((double)(n_fft + 1) * mel_to_hz(mel) / (double)sr);
    }
}

inline void build_filterbank(int n_mels, int n_fft, int sr, double f_min, double f_max, HostTables &t)
{
    const int bins = n_fft / 2 + 1;
    std::vector<int64_t> e;
    mel_bin_edges(n_mels, n_fft, sr, f_min, f_max, e);
    t.fbank.assign((size_t)n_mels * bins, 0.0);
    for (int m = 1; m <= n_mels; m++) {
        const int64_t left = e[m - 1], center = e[m], right = e[m + 1];
        if (right <= left) continue;                       // degenerate filter stays all-zero
        const double up = (double)std::max<int64_t>(1, center - left);
        const double dn = (double)std::max<int64_t>(1, right - center);
        double *row = &t.fbank[(size_t)(m - 1) * bins];
        for (int64_t k = std::max<int64_t>(left, 0); k < std::min<int64_t>(center, bins); k++)
            row[k] = (double)(k - left) / up;
        for (int64_t k = std::max<int64_t>(center, 0); k < std::min<int64_t>(right, bins); k++)
            row[k] = (double)(right - k) / dn;
    }
    // row view
    t.fb_start.assign(n_mels, 0);
    t.fb_cnt.assign(n_mels, 0);
    t.fb_off.assign(n_mels, 0);
    t.fb_w.clear();
    for (int m = 0; m < n_mels; m++) {
        const double *row = &t.fbank[(size_t)m * bins];
        int first = -1, last = -1;
        for (int k = 0; k < bins; k++)
            if (row[k] != 0.0) { if (first < 0) first = k; last = k; }
        t.fb_off[m] = (int32_t)t.fb_w.size();
        if (first >= 0) {
            t.fb_start[m] = first;
            t.fb_cnt[m] = last - first + 1;
            for (int k = first; k <= last; k++) t.fb_w.push_back((float)row[k]);
        }
    }
    // bin view: valid when every bin has at most two non-zero filters and they are adjacent
    t.bin_filt.assign(bins, -2);
    t.bin_wfall.assign(bins, 0.f);
    t.bin_wrise.assign(bins, 0.f);
    t.two_band_ok = true;
    for (int k = 0; k < bins; k++) {
        int nz[3], n = 0;
        for (int m = 0; m < n_mels && n < 3; m++)
            if (t.fbank[(size_t)m * bins + k] != 0.0) nz[n++] = m;
        if (n == 0) continue;
        if (n == 1) {
            // a lone weight is stored as the "fall" slot of its own filter
            t.bin_filt[k] = nz[0];
            t.bin_wfall[k] = (float)t.fbank[(size_t)nz[0] * bins + k];
        } else if (n == 2 && nz[1] == nz[0] + 1) {
            t.bin_filt[k] = nz[0];
            t.bin_wfall[k] = (float)t.fbank[(size_t)nz[0] * bins + k];
            t.bin_wrise[k] = (float)t.fbank[(size_t)nz[1] * bins + k];
        } else {
            t.two_band_ok = false;
        }
    }
    // the warp8 kernel walks runs of equal bin_filt and expects one run per filter: the index must not decrease
    int prev = -1;
    for (int k = 0; k < bins; k++) {
        if (t.bin_filt[k] < 0) continue;
        if (t.bin_filt[k] < prev) t.two_band_ok = false;
        prev = t.bin_filt[k];
    }
}

inline void build_dct2(int n_mfcc, int n_mels, std::vector<double> &d)
{
    d.resize((size_t)n_mfcc * n_mels);
    for (int k = 0; k < n_mfcc; k++)
        for (int n = 0; n < n_mels; n++)
            d[(size_t)k * n_mels + n] = 2.0 * std::cos(M_PI / (double)n_mels * ((double)n + 0.5) * (double)k);
}

}  // namespace dspx
