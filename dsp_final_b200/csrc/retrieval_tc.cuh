// retrieval_tc.cuh -- exact cosine top-k with a tensor-core (tcgen05, TF32) candidate filter.
//
// Same contract as cosine_topk_kernel (retrieval.cuh): float64 scores accumulated left to right with fused
// multiply-adds, descending order, ties to the lower database index -- the result is bit-identical.  The
// similarity matrix q . db^T is a GEMM, so the 5th-generation tensor cores produce an approximate score for
// every pair; a streaming kernel keeps, per query, the k + TC_MARGIN best rows by that score, and a second
// small kernel re-scores only those in float64 and orders them exactly.  Replaces the argsort over the dense
// matrix of src/retrieval/retrieval.py:25-49.
//
// Arithmetic of the filter: kind::tf32 keeps 10 mantissa bits of each operand (~1e-3 on a unit-vector dot
// product).  Each float32 value v is therefore split into hi = v with the low 13 mantissa bits cleared (exact
// in TF32) and lo = v - hi, and the product is accumulated as A_lo B_hi + A_hi B_lo + A_hi B_hi: three K = 32
// passes into the same float32 accumulator.  What is dropped is below 3 * 2^-20 of sum |q_i d_i| <= 1; measured
// against float64 on unit vectors (benchmarks/micro/umma_tf32.cu) the error is 5e-7.  TC_EPS = 2e-5 bounds
// |s_tc - s64| with a factor 40 to spare.
//
// Why the result is exact.  Rows are ordered by (s_tc descending, index ascending).  A member x of the true
// top-k has s64(x) >= s64_k (the exact k-th score), hence s_tc(x) > s64_k - eps.  x can only be missing from
// its split's list if that list is full and its worst entry t_i satisfies t_i >= s_tc(x), or if it was pruned
// against another split's published worst entry t_j > s_tc(x).  The finalize kernel therefore checks
// max_i t_i < s64_k - eps over the full lists: when it holds, every true member was kept and the float64
// re-score orders them exactly; when it does not (more than TC_MARGIN rows within 2 eps of the k-th score,
// e.g. many duplicated rows) the query is flagged and re-ranked by the all-float64 kernel.
//
// Structure of a CTA (352 threads, one per SM), 256 queries x one database split:
//   warps 0-7  epilogue: thread = one query.  tcgen05.ld its accumulator row, compare with the thread's
//              threshold (3-input max tree over 32 columns, bit mask only on a hit), release the TMEM buffer,
//              queue (row, score) in a per-query ring and insert queued pairs into the thread's own unordered
//              list in shared memory (overwrite the worst entry, rescan for the new worst).  No float64 here.
//   warps 8-9  producers: database tile (128 rows x 32 floats, zero padded) from global memory, split into
//              hi / lo, stored K-major under the 128-byte swizzle the tensor core expects; mbarrier hand-off.
//   warp 10    one lane issues 2 x 12 tcgen05.mma (M 128, N 128, K 8) per tile into a double-buffered
//              512-column TMEM accumulator and commits to the mbarriers of the smem stage and the buffer.
#pragma once

#include "retrieval.cuh"

namespace dspx {

#if defined(__CUDACC__)

constexpr int TC_QT = 256;            // queries per CTA: two 128-row A tiles
constexpr int TC_ROWS = 128;          // database rows per tile (UMMA N)
constexpr int TC_STAGES = 2;
constexpr int TC_EPI_THREADS = 256, TC_PROD_THREADS = 64;
constexpr int TC_THREADS = TC_EPI_THREADS + TC_PROD_THREADS + 32;
constexpr int TC_KPAD = 32;           // floats per row of the padded float copies (one 128-byte swizzle row)
constexpr double TC_EPS = 2e-5;
constexpr int TC_FIFO = 8, TC_FIFO_TRIGGER = 4;   // pending candidates per query: ring size / drain trigger
constexpr int TC_MARGIN = 8;          // extra list entries beyond k kept by the approximate score
constexpr int TC_MAX_K = 24;          // the per-thread lists ([k + margin][256] float + int) share smem with the tiles
constexpr int TC_MAX_SPLITS = 16;

template <typename T>
__global__ void normalize_rows_pad32_kernel(const T *x, int64_t n, int dim, double *out64, float *out32)
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
    float *t = out32 + (size_t)r * TC_KPAD;
    for (int c = 0; c < dim; c++) {
        const double v = (double)row[c] / inv;
        out64[(size_t)r * dim + c] = v;
        t[c] = (float)v;
    }
}

#ifdef DSPX_TC_PROFILE
// per-role cycle counters of CTA (0, 0): [0] mma wait acc_empty [1] mma wait full_b [2] mma issue [3] producer wait
// [4] producer store [5] epilogue wait acc_full [6] epilogue masks [7] epilogue enqueue+drain [8] tiles [9] total
__device__ long long tc_prof[16];
#define TC_PROF_T0() long long _pt = clock64()
#define TC_PROF_ADD(i) do { const long long _n = clock64(); if (blockIdx.x == 0 && blockIdx.y == 0) tc_prof_local[i] += _n - _pt; _pt = _n; } while (0)
#else
#define TC_PROF_T0()
#define TC_PROF_ADD(i)
#endif

struct TopkTcParams {
    TopkParams base;        // idx_out / score_out unused here: the lists go to cand_idx / cand_score
    const float *qf;        // [ceil(nq / 256) * 256][32], zero padded
    const float *dbf;       // [ceil(ndb / 128) * 128][32], zero padded
    unsigned int *shared_thr;   // [nq] order-encoded worst list scores shared by the splits of a query (null: one split)
    int kp;                 // list length: k + TC_MARGIN
    int32_t *cand_idx;      // [nq][n_splits][kp] rows kept (-1: empty), unordered
    float *cand_score;      // [nq][n_splits][kp] their tensor-core scores
};

// order-preserving float <-> uint32 (0 is below every float): lets atomicMax maintain a shared threshold
__device__ __forceinline__ unsigned int tc_enc(float f)
{
    const unsigned int b = __float_as_uint(f);
    return (b >> 31) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float tc_dec(unsigned int u)
{
    return __uint_as_float((u >> 31) ? (u & 0x7fffffffu) : ~u);
}
// largest float below f (f finite)
__device__ __forceinline__ float tc_below(float f) { return tc_dec(tc_enc(f) - 1u); }

__device__ __forceinline__ uint32_t tc_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tc_mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t *bar)
{
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(tc_smem_u32(bar)) : "memory");
}
// bounded wait: a protocol error traps (launch failure) instead of hanging the GPU
__device__ __forceinline__ void tc_mbar_wait(uint64_t *bar, uint32_t parity)
{
    for (uint32_t spin = 0;; spin++) {
        uint32_t ok;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(tc_smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
        if (spin > (1u << 28)) __trap();
    }
}

// K-major operand tile under SWIZZLE_128B: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t tc_desc(const void *tile)
{
    return (uint64_t)((tc_smem_u32(tile) >> 4) & 0x3fff) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}

__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
                 ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}

// split a float4 into TF32-exact high parts and residuals; store both at the swizzled chunk position
__device__ __forceinline__ void tc_store_split(unsigned char *hi_tile, unsigned char *lo_tile, int row, int chunk, float4 v)
{
    const uint32_t off = (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
    float4 h;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
    *reinterpret_cast<float4 *>(hi_tile + off) = h;
    *reinterpret_cast<float4 *>(lo_tile + off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
}

// 32 accumulator columns of this thread's TMEM lane: issue, and wait.  The wait names the destination registers as
// in/out operands so that no use of them can be scheduled above it.
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait(uint32_t (&v)[32])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                   "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :: "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    tc_ld32_issue(taddr, v);
    tc_ld_wait(v);
}
// v[j] for a run-time j: five levels of selects (register arrays cannot be indexed dynamically)
__device__ __forceinline__ float tc_pick(const uint32_t *w, int j)
{
    uint32_t a[16], b[8], c[4];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = (j & 1) ? w[2 * i + 1] : w[2 * i];
#pragma unroll
    for (int i = 0; i < 8; i++) b[i] = (j & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
    for (int i = 0; i < 4; i++) c[i] = (j & 4) ? b[2 * i + 1] : b[2 * i];
    const uint32_t d0 = (j & 8) ? c[1] : c[0], d1 = (j & 8) ? c[3] : c[2];
    return __uint_as_float((j & 16) ? d1 : d0);
}

inline size_t topk_tc_smem_bytes(int kp)
{
    return (size_t)(4 + 2 * TC_STAGES) * TC_ROWS * 128 + (size_t)TC_QT * kp * 8 + (size_t)TC_FIFO * TC_QT * 8 + 128;
}

__global__ void __launch_bounds__(TC_THREADS, 1) cosine_topk_tc_kernel(const TopkTcParams pp)
{
    const TopkParams &p = pp.base;
    extern __shared__ __align__(1024) unsigned char tc_raw[];     // the 128-byte swizzle pattern repeats every 1024 bytes
    if (tc_smem_u32(tc_raw) & 1023u) __trap();
    unsigned char *a_hi = tc_raw;                                 // [2][128 rows x 128 B]
    unsigned char *a_lo = a_hi + 2 * TC_ROWS * 128;
    unsigned char *b_hi = a_lo + 2 * TC_ROWS * 128;               // [TC_STAGES][128 x 128 B]
    unsigned char *b_lo = b_hi + TC_STAGES * TC_ROWS * 128;
    float *s_ls = reinterpret_cast<float *>(b_lo + TC_STAGES * TC_ROWS * 128);       // [kp][256] list scores
    int32_t *s_li = reinterpret_cast<int32_t *>(s_ls + (size_t)pp.kp * TC_QT);       // [kp][256] list rows
    float *s_fs = reinterpret_cast<float *>(s_li + (size_t)pp.kp * TC_QT);           // [TC_FIFO][256] pending scores
    int32_t *s_fi = reinterpret_cast<int32_t *>(s_fs + TC_FIFO * TC_QT);             // [TC_FIFO][256] pending rows
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_fi + TC_FIFO * TC_QT);           // 8-byte aligned (all sizes are multiples of 1 KB)
    uint64_t *full_b = bars, *empty_b = bars + TC_STAGES, *acc_full = bars + 2 * TC_STAGES, *acc_empty = acc_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t q0 = (int64_t)blockIdx.x * TC_QT;
    const int split = blockIdx.y;
    const int64_t r_begin = (int64_t)split * p.rows_per_split;    // multiple of TC_ROWS
    int64_t r_end = r_begin + p.rows_per_split;
    if (r_end > p.ndb) r_end = p.ndb;
    const int64_t n_tiles = (r_end - r_begin + TC_ROWS - 1) / TC_ROWS;

    // ---- set-up: query tiles (split, swizzled), barriers, TMEM ----------------------------------
    for (int c = tid; c < TC_QT * 8; c += TC_THREADS) {
        const int row = c >> 3, chunk = c & 7, a = row >> 7;
        const float4 v = *reinterpret_cast<const float4 *>(pp.qf + ((size_t)(q0 + row) * TC_KPAD + chunk * 4));
        tc_store_split(a_hi + a * TC_ROWS * 128, a_lo + a * TC_ROWS * 128, row & 127, chunk, v);
    }
    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; s++) { tc_mbar_init(&full_b[s], TC_PROD_THREADS); tc_mbar_init(&empty_b[s], 1); }
        for (int b = 0; b < 2; b++) { tc_mbar_init(&acc_full[b], 1); tc_mbar_init(&acc_empty[b], TC_EPI_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == TC_THREADS / 32 - 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tc_smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");               // the A tiles were written through the generic proxy
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *tmem_slot;

    if (tid < TC_EPI_THREADS) {
        // ===== epilogue: one query per thread =====
        const int ql = tid;                                        // = (warp >> 2) * 128 + (warp & 3) * 32 + lane
        const int64_t gq = q0 + ql;
        const bool active = gq < p.nq;
        float *ls = s_ls + ql;                                     // entry e at ls[e * TC_QT]
        int32_t *li = s_li + ql;
        float *fs = s_fs + ql;
        int32_t *fi = s_fi + ql;
        const int kp = pp.kp;
        int cnt = 0, wpos = 0;                                     // list fill, position of its worst entry
        float thr = -INFINITY;                                     // score of the worst entry once the list is full
        int f_head = 0, f_cnt = 0;
        // Splits of one query share the best "worst kept score" any of them has reached: a row scoring below it is
        // beaten by k + margin rows of that split and can be skipped.  This only prunes work.
        unsigned int *gslot = (pp.shared_thr && active) ? pp.shared_thr + gq : nullptr;
        float gthr = -INFINITY;
        unsigned int genc = 0;
#ifdef DSPX_TC_PROFILE
        long long tc_prof_local[16] = {0};
#endif
        // pending (row, score) pairs -> list.  Rows arrive in increasing order, so a pair that ties the worst
        // entry loses (strict >), and the worst entry is the lowest score with the highest row.
        auto drain = [&]() {
            while (__any_sync(0xffffffffu, f_cnt > 0)) {
                if (f_cnt == 0) continue;
                const int slot = f_head & (TC_FIFO - 1);
                const float s = fs[(size_t)slot * TC_QT];
                const int32_t row = fi[(size_t)slot * TC_QT];
                f_head++;
                f_cnt--;
                if (cnt == kp && !(s > thr)) continue;
                const int pos = cnt < kp ? cnt : wpos;
                ls[(size_t)pos * TC_QT] = s;
                li[(size_t)pos * TC_QT] = row;
                if (cnt < kp) cnt++;
                if (cnt == kp) {
                    float w0 = INFINITY, w1 = INFINITY;
                    int32_t i0 = -1, i1 = -1;
                    int p0 = 0, p1 = 0;
                    int e = 0;
                    for (; e + 1 < kp; e += 2) {                    // two independent scans
                        const float a = ls[(size_t)e * TC_QT], b = ls[(size_t)(e + 1) * TC_QT];
                        const int32_t ai = li[(size_t)e * TC_QT], bi = li[(size_t)(e + 1) * TC_QT];
                        if (a < w0 || (a == w0 && ai > i0)) { w0 = a; i0 = ai; p0 = e; }
                        if (b < w1 || (b == w1 && bi > i1)) { w1 = b; i1 = bi; p1 = e + 1; }
                    }
                    if (e < kp) {
                        const float a = ls[(size_t)e * TC_QT];
                        const int32_t ai = li[(size_t)e * TC_QT];
                        if (a < w0 || (a == w0 && ai > i0)) { w0 = a; i0 = ai; p0 = e; }
                    }
                    if (w1 < w0 || (w1 == w0 && i1 > i0)) { w0 = w1; p0 = p1; }
                    const bool rose = w0 > thr;
                    thr = w0;
                    wpos = p0;
                    if (gslot && rose && thr > gthr) atomicMax(gslot, tc_enc(thr));
                }
            }
        };
        // one block of 32 accumulator columns: max tree first (almost every block is entirely below the threshold),
        // bit mask and queueing only when some lane of the warp has a hit
        auto scan32 = [&](const uint32_t *w, int64_t row0, float T) {
            float mx[11];
#pragma unroll
            for (int j = 0; j < 10; j++)
                mx[j] = fmaxf(fmaxf(__uint_as_float(w[3 * j]), __uint_as_float(w[3 * j + 1])), __uint_as_float(w[3 * j + 2]));
            mx[10] = fmaxf(__uint_as_float(w[30]), __uint_as_float(w[31]));
            const float t0 = fmaxf(fmaxf(mx[0], mx[1]), mx[2]), t1 = fmaxf(fmaxf(mx[3], mx[4]), mx[5]);
            const float t2 = fmaxf(fmaxf(mx[6], mx[7]), mx[8]), t3 = fmaxf(mx[9], mx[10]);
            const float top = fmaxf(fmaxf(t0, t1), fmaxf(t2, t3));
            if (!__any_sync(0xffffffffu, top > T)) return;
            uint32_t part[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int j = 0; j < 32; j++) part[j & 3] |= (__uint_as_float(w[j]) > T) ? (1u << j) : 0u;
            uint32_t m = (part[0] | part[1]) | (part[2] | part[3]);
            if (row0 + 32 > r_end) {                               // rows past the end of the split score 0: never candidates
                const int64_t valid = r_end - row0;
                m = valid <= 0 ? 0u : (m & (0xffffffffu >> (32 - (int)valid)));
            }
            while (__any_sync(0xffffffffu, m != 0)) {
                if (m != 0 && f_cnt < TC_FIFO) {
                    const int j = __ffs(m) - 1;
                    m &= m - 1;
                    const int slot = (f_head + f_cnt) & (TC_FIFO - 1);
                    fs[(size_t)slot * TC_QT] = tc_pick(w, j);
                    fi[(size_t)slot * TC_QT] = (int32_t)(row0 + j);
                    f_cnt++;
                }
                if (__any_sync(0xffffffffu, m != 0 && f_cnt == TC_FIFO)) drain();
            }
        };
        auto filter_threshold = [&]() -> float {                   // a row is a candidate when its score is > this
            if (!active) return INFINITY;
            const float g = gthr == -INFINITY ? -INFINITY : tc_below(gthr);      // ties with another split's worst stay in
            return cnt == kp ? fmaxf(thr, g) : g;
        };
        if (gslot) genc = __ldcg(gslot);                           // what earlier CTAs of this query already reached
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * TC_ROWS);
        TC_PROF_T0();
        for (int64_t t = 0; t < n_tiles; t++) {
            const int buf = (int)(t & 1);
            const int64_t tile = r_begin + t * TC_ROWS;
            tc_mbar_wait(&acc_full[buf], (uint32_t)((t >> 1) & 1));
            __syncwarp();                                          // tcgen05.ld is warp-collective
            asm volatile("tcgen05.fence::after_thread_sync;");
            TC_PROF_ADD(5);
            if (genc) gthr = tc_dec(genc);
            const uint32_t acc = lane_base + (uint32_t)(buf * 2 * TC_ROWS);
            if (t == 0) {
                // the list is empty: take the first tile 32 columns at a time so the threshold tightens as it fills
#pragma unroll 1
                for (int cb = 0; cb < 4; cb++) {
                    uint32_t v[32];
                    tc_ld32(acc + cb * 32, v);
                    scan32(v, tile + cb * 32, filter_threshold());
                    drain();
                }
                asm volatile("tcgen05.fence::before_thread_sync;");
                __syncwarp();
                if (lane == 0) tc_mbar_arrive(&acc_empty[buf]);
            } else {
                // two register buffers: the load of block i + 1 is in flight while block i is scanned
                const float T = filter_threshold();
                uint32_t va[32], vb[32];
                tc_ld32_issue(acc, va);
                tc_ld_wait(va);
                tc_ld32_issue(acc + 32, vb);
                scan32(va, tile, T);
                tc_ld_wait(vb);
                tc_ld32_issue(acc + 64, va);
                scan32(vb, tile + 32, T);
                tc_ld_wait(va);
                tc_ld32_issue(acc + 96, vb);
                scan32(va, tile + 64, T);
                tc_ld_wait(vb);
                asm volatile("tcgen05.fence::before_thread_sync;");
                __syncwarp();
                if (lane == 0) tc_mbar_arrive(&acc_empty[buf]);        // the tensor core may overwrite this buffer now
                scan32(vb, tile + 96, T);
            }
            if (gslot) genc = __ldcg(gslot);                       // for the next tile
            TC_PROF_ADD(6);
            if (t < 4 || t == n_tiles - 1 || __any_sync(0xffffffffu, f_cnt >= TC_FIFO_TRIGGER)) drain();
            TC_PROF_ADD(7);
        }
#ifdef DSPX_TC_PROFILE
        if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) { for (int i = 5; i < 8; i++) tc_prof[i] = tc_prof_local[i]; tc_prof[8] = n_tiles; }
#endif
        if (active) {
            const size_t o = ((size_t)gq * p.n_splits + split) * kp;
            for (int e = 0; e < kp; e++) {
                const bool have = e < cnt;
                pp.cand_idx[o + e] = have ? li[(size_t)e * TC_QT] : -1;
                pp.cand_score[o + e] = have ? ls[(size_t)e * TC_QT] : -INFINITY;
            }
        }
    } else if (tid < TC_EPI_THREADS + TC_PROD_THREADS) {
        // ===== producers: database tile -> hi / lo operand tiles =====
        // A tile is 16 KB and a load takes ~1 us to return: every thread keeps its 16 float4 of the NEXT tile in
        // flight while it waits for the stage to be released (Little's law: less than a tile in flight starves
        // the tensor core).
        const int ptid = tid - TC_EPI_THREADS;
        constexpr int PER = TC_ROWS * 8 / TC_PROD_THREADS;         // 16 chunks of 16 bytes per thread
        float4 v[PER];
        auto fetch = [&](int64_t t) {
            const float4 *src = reinterpret_cast<const float4 *>(pp.dbf + (size_t)(r_begin + t * TC_ROWS) * TC_KPAD);
#pragma unroll
            for (int i = 0; i < PER; i++) v[i] = __ldg(src + ptid + i * TC_PROD_THREADS);
        };
        if (n_tiles > 0) fetch(0);
#ifdef DSPX_TC_PROFILE
        long long tc_prof_local[16] = {0};
#endif
        TC_PROF_T0();
        for (int64_t t = 0; t < n_tiles; t++) {
            const int s = (int)(t % TC_STAGES);
            tc_mbar_wait(&empty_b[s], (uint32_t)(((t / TC_STAGES) & 1) ^ 1));
            TC_PROF_ADD(3);
            unsigned char *hi = b_hi + s * TC_ROWS * 128, *lo = b_lo + s * TC_ROWS * 128;
#pragma unroll
            for (int i = 0; i < PER; i++) {
                const int c = ptid + i * TC_PROD_THREADS;
                tc_store_split(hi, lo, c >> 3, c & 7, v[i]);
            }
            asm volatile("fence.proxy.async.shared::cta;");
            tc_mbar_arrive(&full_b[s]);
            if (t + 1 < n_tiles) fetch(t + 1);
            TC_PROF_ADD(4);
        }
#ifdef DSPX_TC_PROFILE
        if (ptid == 0 && blockIdx.x == 0 && blockIdx.y == 0) { tc_prof[3] = tc_prof_local[3]; tc_prof[4] = tc_prof_local[4]; }
#endif
    } else if (lane == 0) {
        // ===== tensor-core issue: one thread =====
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_ROWS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
#ifdef DSPX_TC_PROFILE
        long long tc_prof_local[16] = {0};
        const long long tc_start = clock64();
#endif
        TC_PROF_T0();
        for (int64_t t = 0; t < n_tiles; t++) {
            const int s = (int)(t % TC_STAGES), buf = (int)(t & 1);
            tc_mbar_wait(&acc_empty[buf], (uint32_t)(((t >> 1) & 1) ^ 1));
            TC_PROF_ADD(0);
            tc_mbar_wait(&full_b[s], (uint32_t)((t / TC_STAGES) & 1));
            TC_PROF_ADD(1);
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint64_t bh = tc_desc(b_hi + s * TC_ROWS * 128), bl = tc_desc(b_lo + s * TC_ROWS * 128);
#pragma unroll
            for (int a = 0; a < 2; a++) {
                const uint64_t ah = tc_desc(a_hi + a * TC_ROWS * 128), al = tc_desc(a_lo + a * TC_ROWS * 128);
                const uint32_t d = tmem + (uint32_t)(buf * 2 * TC_ROWS + a * TC_ROWS);
#pragma unroll
                for (int kk = 0; kk < 4; kk++) tc_mma_tf32(d, al + 2 * kk, bh + 2 * kk, idesc, kk > 0);     // small terms first
#pragma unroll
                for (int kk = 0; kk < 4; kk++) tc_mma_tf32(d, ah + 2 * kk, bl + 2 * kk, idesc, 1);
#pragma unroll
                for (int kk = 0; kk < 4; kk++) tc_mma_tf32(d, ah + 2 * kk, bh + 2 * kk, idesc, 1);
            }
            tc_commit(&empty_b[s]);                                // smem stage free once these MMAs have read it
            tc_commit(&acc_full[buf]);                             // accumulators complete
            TC_PROF_ADD(2);
        }
#ifdef DSPX_TC_PROFILE
        if (blockIdx.x == 0 && blockIdx.y == 0) { for (int i = 0; i < 3; i++) tc_prof[i] = tc_prof_local[i]; tc_prof[9] = clock64() - tc_start; }
#endif
    }

    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    __syncwarp();
    if (warp == TC_THREADS / 32 - 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// ---- finalize: exact float64 re-score of the kept rows, exact order, and the safety check -----------------
// One warp per query.  Candidates: n_splits lists of kp (row, s_tc) pairs.
constexpr int TC_FIN_WARPS = 4;
constexpr int TC_FIN_MAXC = TC_MAX_SPLITS * (TC_MAX_K + TC_MARGIN);

__global__ void __launch_bounds__(TC_FIN_WARPS * 32) topk_tc_finalize_kernel(const int32_t *cand_idx, const float *cand_score,
                                                                            int n_splits, int kp, const double *qn,
                                                                            const double *dbn, int64_t nq, int dim, int k,
                                                                            int32_t *idx_out, double *score_out,
                                                                            unsigned char *flags, unsigned int *n_flagged)
{
    __shared__ double s_sc[TC_FIN_WARPS][TC_FIN_MAXC];
    __shared__ int32_t s_ix[TC_FIN_WARPS][TC_FIN_MAXC];
    __shared__ double s_q[TC_FIN_WARPS][32];
    __shared__ float s_cs[TC_FIN_WARPS][TC_FIN_MAXC];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t q = (int64_t)blockIdx.x * TC_FIN_WARPS + w;
    if (q >= nq) return;
    const int C = n_splits * kp;
    const int32_t *ci = cand_idx + (size_t)q * C;
    const float *cs = cand_score + (size_t)q * C;
    if (lane < dim) s_q[w][lane] = qn[(size_t)q * dim + lane];
    for (int e = lane; e < C; e += 32) {                    // stage the lists: coalesced
        s_ix[w][e] = ci[e];
        s_cs[w][e] = cs[e];
    }
    __syncwarp();
    // worst tensor-core score of every full list (lane = split)
    float tmax = -INFINITY;
    if (lane < n_splits) {
        float t = INFINITY;
        int32_t all = 0;
        for (int e = 0; e < kp; e++) {
            all |= s_ix[w][lane * kp + e];                  // negative (empty slot) sets the sign bit
            t = fminf(t, s_cs[w][lane * kp + e]);
        }
        if (all >= 0) tmax = t;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, off));
    __syncwarp();
    // A full list holds k + margin rows scoring >= tmax, so a row scoring below tmax - 2 eps is beaten exactly by
    // at least k of them: only the rows above that line are re-scored (compacted with ballots).
    const float keep = tmax == -INFINITY ? -INFINITY : (float)((double)tmax - 2.0 * TC_EPS) - 1e-6f;
    int n_kept = 0;
    for (int e0 = 0; e0 < C; e0 += 32) {
        const int e = e0 + lane;
        const int32_t row = e < C ? s_ix[w][e] : -1;
        const bool take = row >= 0 && s_cs[w][e] >= keep;
        const unsigned bal = __ballot_sync(0xffffffffu, take);   // every lane has read its entry: slots <= e may be rewritten
        if (take) {
            const double *d = dbn + (size_t)row * dim;
            double dv[32];                                  // dim <= 32: all loads in flight before the chain starts
#pragma unroll
            for (int c = 0; c < 32; c++) dv[c] = c < dim ? d[c] : 0.0;
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < 32; c++)
                if (c < dim) s = fma(s_q[w][c], dv[c], s);                 // the oracle's chain
            const int slot = n_kept + __popc(bal & ((1u << lane) - 1u));
            s_sc[w][slot] = s;
            s_ix[w][slot] = row;
        }
        n_kept += __popc(bal);
    }
    __syncwarp();
    // k rounds of arg-best under (score descending, row ascending)
    double kth = 0.0;
    for (int out = 0; out < k; out++) {
        double bs = -INFINITY;
        int32_t bi = 0x7fffffff;
        int be = -1;
        for (int e = lane; e < n_kept; e += 32) {
            const double s = s_sc[w][e];
            const int32_t i = s_ix[w][e];
            if (i >= 0 && (s > bs || (s == bs && i < bi))) { bs = s; bi = i; be = e; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, bs, off);
            const int32_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
            const int oe = __shfl_xor_sync(0xffffffffu, be, off);
            if (os > bs || (os == bs && oi < bi)) { bs = os; bi = oi; be = oe; }
        }
        if (lane == 0) {
            idx_out[(size_t)q * k + out] = be >= 0 ? bi : -1;
            if (score_out) score_out[(size_t)q * k + out] = bs;
            if (be >= 0) s_ix[w][be] = -1;
        }
        kth = bs;
        __syncwarp();
    }
    if (lane == 0) {
        const bool unsafe = (double)tmax >= kth - TC_EPS;          // a list may have dropped a member of the top-k
        flags[q] = unsafe ? 1 : 0;
        if (unsafe) atomicAdd(n_flagged, 1u);
    }
}


#endif

}  // namespace dspx
