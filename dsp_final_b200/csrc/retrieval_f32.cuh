// retrieval_f32.cuh -- exact cosine top-k with an FP32 pre-filter.
//
// Same contract as cosine_topk_kernel (retrieval.cuh): float64 scores accumulated left to right with
// fused multiply-adds, descending order, ties to the lower database index -- the result is bit-identical.
// The difference is where the time goes.  Almost every (query, row) pair is far below the query's current
// k-th score, so pairs are first scored in FP32 with packed FFMA2 on float copies of the normalised
// vectors; only pairs whose FP32 score comes within TK_EPS of the k-th float64 score are re-scored
// exactly in float64 and offered to the running top-k list.
//
// Why this is exact: rows are unit vectors, so |q . d| <= 1 and sum |q_i d_i| <= 1.  Rounding both
// vectors to float perturbs the dot product by < 1.3e-7; accumulating 'dim' float FMAs adds at most
// dim * 2^-24 (2.4e-6 for dim 26, 6.2e-5 for dim 1024).  TK_EPS below is scaled with dim so that
// |s32 - s64| < TK_EPS always holds; hence every pair with s64 >= thr64 has s32 > thr32 = float(thr64) - eps
// and is re-scored.  False positives only cost a few extra float64 dot products.
//
// Data layout: the normalise kernel also writes tile-transposed float copies, [tile][dim][rows_per_tile]
// (128 database rows or 128 queries per tile, zero padded), so a tile is one contiguous block:
// 16-byte cp.async in, conflict-free 128-bit shared loads out (lane l owns rows 4l..4l+3 of the tile).
#pragma once

#include "retrieval.cuh"

namespace dspx {

#if defined(__CUDACC__)

template <typename T>
__global__ void normalize_rows_tiled_kernel(const T *x, int64_t n, int dim, double *out64, float *out32t, int tile_rows)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const T *row = x + (size_t)r * dim;
    double s = 0.0;
    for (int c = 0; c < dim; c++) {
        const double v = (double)row[c];
        s = fma(v, v, s);
    }
    const double inv = sqrt(s) + 1e-10;
    float *t = out32t + (size_t)(r / tile_rows) * dim * tile_rows + (r % tile_rows);
    for (int c = 0; c < dim; c++) {
        const double v = (double)row[c] / inv;
        out64[(size_t)r * dim + c] = v;
        t[(size_t)c * tile_rows] = (float)v;
    }
}

struct TopkF32Params {
    TopkParams base;
    const float *qf_t;      // [ceil(nq/128)][dim][128]
    const float *dbf_t;     // [ceil(ndb/128)][dim][128]
    float eps;
};

__device__ __forceinline__ bool tk_better(double s, int32_t i, double s2, int32_t i2) { return s > s2 || (s == s2 && i < i2); }

// sorted insert with the full (score desc, index asc) comparator: arrival order is arbitrary here
__device__ __forceinline__ void topk_warp_insert_any(double *ls, int32_t *li, int k, int &cnt, double cs, int32_t ci, int lane)
{
    int p = 0;
    for (int base = 0; base < cnt; base += 32) {
        const int e = base + lane;
        const bool ahead = (e < cnt) && tk_better(ls[e], li[e], cs, ci);
        p += __popc(__ballot_sync(0xffffffffu, ahead));
    }
    const int last = cnt < k ? cnt : k - 1;
    if (p > last) return;
    for (int base = ((last - 1) >= 0 ? (last - 1) / 32 : 0) * 32; base >= 0 && base + 32 > p; base -= 32) {
        const int e = base + lane;
        const bool mv = (e >= p) && (e < last);
        double vs = 0.0;
        int32_t vi = 0;
        if (mv) { vs = ls[e]; vi = li[e]; }
        __syncwarp();
        if (mv) { ls[e + 1] = vs; li[e + 1] = vi; }
        __syncwarp();
    }
    if (lane == 0) { ls[p] = cs; li[p] = ci; }
    if (cnt < k) cnt++;
    __syncwarp();
}

constexpr int TKF_WARPS = 16;                    // one 16-warp CTA per SM: 128 queries share every database tile
constexpr int TKF_QPC = TKF_WARPS * TK_QPW;      // 128

// slow path: exact float64 re-score of this lane's rows that passed the FP32 filter -- from the float64 copy
// of the same tile in shared memory -- then ordered offers to the running top-k list.
__device__ __noinline__ TopkState topk_offer_f32(float s0, float s1, float s2, float s3, int64_t tile, int64_t r_end,
                                                 const double *q64, const double *db64, int dim, float eps, double *ls,
                                                 int32_t *li, int k, int cnt, double thr, int lane)
{
    const float thr32 = cnt < k ? -INFINITY : (float)thr - eps;
#pragma unroll 1
    for (int r = 0; r < TK_RPL; r++) {
        const float s = r == 0 ? s0 : (r == 1 ? s1 : (r == 2 ? s2 : s3));
        const int64_t gr = tile + 4 * lane + r;
        const bool take = (gr < r_end) && (s > thr32);
        double s64 = 0.0;
        if (take) {
            const double *d = db64 + (size_t)(4 * lane + r) * dim;
            int c = 0;
            for (; c + 8 <= dim; c += 8) {                 // loads first, then the ordered fma chain
                double qv[8], dv[8];
#pragma unroll
                for (int j = 0; j < 8; j++) { qv[j] = q64[c + j]; dv[j] = d[c + j]; }
#pragma unroll
                for (int j = 0; j < 8; j++) s64 = fma(qv[j], dv[j], s64);
            }
            for (; c < dim; c++) s64 = fma(q64[c], d[c], s64);
        }
        unsigned mask = __ballot_sync(0xffffffffu, take);
        while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const double cs = __shfl_sync(0xffffffffu, s64, src);
            const int32_t ci = (int32_t)(tile + 4 * src + r);
            if (cnt == k && !tk_better(cs, ci, ls[k - 1], li[k - 1])) continue;
            topk_warp_insert_any(ls, li, k, cnt, cs, ci, lane);
            if (cnt == k) thr = ls[k - 1];
        }
    }
    TopkState out;
    out.thr = thr;
    out.cnt = cnt;
    return out;
}

// Single-chunk kernel (dim <= CW): every stage brings one database tile twice -- transposed floats for the
// FP32 scoring loop and row-major float64 for exact re-scoring -- double-buffered with cp.async.
template <int CW>
__global__ void __launch_bounds__(TKF_WARPS * 32, 1) cosine_topk_f32_kernel(const TopkF32Params pp)
{
    const TopkParams &p = pp.base;
    extern __shared__ __align__(16) unsigned char tkf_raw[];
    const int dim = p.dim;
    double *s_db64 = reinterpret_cast<double *>(tkf_raw);              // [2][TK_ROWS * dim]
    double *s_q64 = s_db64 + 2 * TK_ROWS * dim;                        // [TKF_QPC][dim]
    double *s_ls = s_q64 + TKF_QPC * dim;                              // [TKF_QPC][k]
    float *s_db = reinterpret_cast<float *>(s_ls + (size_t)TKF_QPC * p.k);   // [2][CW][TK_ROWS]
    float *s_q = s_db + 2 * CW * TK_ROWS;                              // [CW][TKF_QPC]
    int32_t *s_li = reinterpret_cast<int32_t *>(s_q + CW * TKF_QPC);   // [TKF_QPC][k]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q0 = (int64_t)blockIdx.x * TKF_QPC;
    const int split = blockIdx.y;
    const int64_t r_begin = (int64_t)split * p.rows_per_split;         // multiple of TK_ROWS
    int64_t r_end = r_begin + p.rows_per_split;
    if (r_end > p.ndb) r_end = p.ndb;
    const int64_t n_tiles = (r_end - r_begin + TK_ROWS - 1) / TK_ROWS;

    auto cp16 = [](void *dst, const void *src, int nbytes) {
        const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(nbytes));
    };
    auto issue = [&](int64_t t) {
        const int buf = (int)(t & 1);
        const int64_t tile = r_begin + t * TK_ROWS;
        const float *src = pp.dbf_t + (size_t)(tile / TK_ROWS) * dim * TK_ROWS;        // [dim][128], zero padded
        float *dst = s_db + buf * CW * TK_ROWS;
        for (int e = tid; e < CW * TK_ROWS / 4; e += TKF_WARPS * 32) {
            const bool ok = e / (TK_ROWS / 4) < dim;
            cp16(dst + 4 * e, src + (ok ? 4 * e : 0), ok ? 16 : 0);
        }
        const int64_t rows = (r_end - tile) < TK_ROWS ? (r_end - tile) : TK_ROWS;
        const int64_t valid_bytes = rows * dim * 8;                                  // tile start is 16-byte aligned
        const char *s64 = reinterpret_cast<const char *>(p.dbn + (size_t)tile * dim);
        char *d64 = reinterpret_cast<char *>(s_db64 + (size_t)buf * TK_ROWS * dim);
        for (int e = tid; e < TK_ROWS * dim / 2; e += TKF_WARPS * 32) {
            const int64_t left = valid_bytes - 16 * (int64_t)e;
            const int nb = left >= 16 ? 16 : (left > 0 ? (int)left : 0);
            cp16(d64 + 16 * e, s64 + (nb ? 16 * (size_t)e : 0), nb);
        }
        if (t == 0) {
            const float *qs = pp.qf_t + (size_t)blockIdx.x * dim * TKF_QPC;          // [dim][128], zero padded
            for (int e = tid; e < CW * TKF_QPC / 4; e += TKF_WARPS * 32) {
                const bool ok = e / (TKF_QPC / 4) < dim;
                cp16(s_q + 4 * e, qs + (ok ? 4 * e : 0), ok ? 16 : 0);
            }
            const int64_t qrows = (p.nq - q0) < TKF_QPC ? (p.nq - q0) : TKF_QPC;
            const int64_t qbytes = qrows * dim * 8;
            const char *q64 = reinterpret_cast<const char *>(p.qn + (size_t)q0 * dim);
            for (int e = tid; e < TKF_QPC * dim / 2; e += TKF_WARPS * 32) {
                const int64_t left = qbytes - 16 * (int64_t)e;
                const int nb = left >= 16 ? 16 : (left > 0 ? (int)left : 0);
                cp16(reinterpret_cast<char *>(s_q64) + 16 * e, q64 + (nb ? 16 * (size_t)e : 0), nb);
            }
        }
        tk_cp_commit();
    };

    int cnt[TK_QPW];
    double thr[TK_QPW];
#pragma unroll
    for (int qi = 0; qi < TK_QPW; qi++) { cnt[qi] = 0; thr[qi] = 0.0; }

    if (n_tiles > 0) issue(0);
    for (int64_t t = 0; t < n_tiles; t++) {
        const int buf = (int)(t & 1);
        const int64_t tile = r_begin + t * TK_ROWS;
        if (t + 1 < n_tiles) { issue(t + 1); tk_cp_wait<1>(); }
        else tk_cp_wait<0>();
        __syncthreads();
        float2 acc[TK_QPW][2];
#pragma unroll
        for (int qi = 0; qi < TK_QPW; qi++) { acc[qi][0] = make_float2(0.f, 0.f); acc[qi][1] = make_float2(0.f, 0.f); }
        const float *db = s_db + buf * CW * TK_ROWS + 4 * lane;
        const float *qq = s_q + warp * TK_QPW;
#pragma unroll
        for (int c = 0; c < CW; c++) {
            const float4 dv = *reinterpret_cast<const float4 *>(db + c * TK_ROWS);
            const float4 qa = *reinterpret_cast<const float4 *>(qq + c * TKF_QPC);
            const float4 qb = *reinterpret_cast<const float4 *>(qq + c * TKF_QPC + 4);
            const float qv[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
            const float2 d01 = make_float2(dv.x, dv.y), d23 = make_float2(dv.z, dv.w);
#pragma unroll
            for (int qi = 0; qi < TK_QPW; qi++) {
                const float2 qv2 = make_float2(qv[qi], qv[qi]);
                acc[qi][0] = __ffma2_rn(d01, qv2, acc[qi][0]);
                acc[qi][1] = __ffma2_rn(d23, qv2, acc[qi][1]);
            }
        }
#pragma unroll
        for (int qi = 0; qi < TK_QPW; qi++) {
            const int ql = warp * TK_QPW + qi;
            if (q0 + ql >= p.nq) continue;                              // warp-uniform
            const float mx = fmaxf(fmaxf(acc[qi][0].x, acc[qi][0].y), fmaxf(acc[qi][1].x, acc[qi][1].y));
            const float thr32 = cnt[qi] < p.k ? -INFINITY : (float)thr[qi] - pp.eps;
            if (__any_sync(0xffffffffu, mx > thr32)) {
                const TopkState st = topk_offer_f32(acc[qi][0].x, acc[qi][0].y, acc[qi][1].x, acc[qi][1].y, tile, r_end,
                                                    s_q64 + (size_t)ql * dim, s_db64 + (size_t)buf * TK_ROWS * dim, dim, pp.eps,
                                                    s_ls + (size_t)ql * p.k, s_li + (size_t)ql * p.k, p.k, cnt[qi], thr[qi], lane);
                cnt[qi] = st.cnt;
                thr[qi] = st.thr;
            }
        }
        __syncthreads();
    }

    __syncwarp();
    for (int qi = 0; qi < TK_QPW; qi++) {
        const int ql = warp * TK_QPW + qi;
        const int64_t gq = q0 + ql;
        if (gq >= p.nq) continue;
        const double *ls = s_ls + (size_t)ql * p.k;
        const int32_t *li = s_li + (size_t)ql * p.k;
        const size_t o = ((size_t)gq * p.n_splits + split) * p.k;
        for (int e = lane; e < p.k; e += 32) {
            const bool have = e < cnt[qi];
            p.idx_out[o + e] = have ? li[e] : -1;
            if (p.score_out) p.score_out[o + e] = have ? ls[e] : -INFINITY;
        }
    }
}

inline size_t topk_f32_smem_bytes(int dim, int k, int cw)
{
    return (size_t)(2 * TK_ROWS * dim + TKF_QPC * dim + TKF_QPC * k) * 8 + (size_t)(2 * cw * TK_ROWS + cw * TKF_QPC) * 4 +
           (size_t)TKF_QPC * k * 4;
}
#endif

}  // namespace dspx
