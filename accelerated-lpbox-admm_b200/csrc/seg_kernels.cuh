// Unconstrained Lp-Box ADMM (graph-cut segmentation, `min x'Ax + b'x`) for sm_100a -- streaming variant.
//
// An image-sized problem (n = 187 500 for 375x500) does not fit on chip: per image 11 fp64 n-vectors + A (<= 7 stored
// entries per row).  One CTA owns one image for a whole window and streams its vectors through L2/HBM with coalesced
// accesses; there is NO inter-CTA synchronisation, so a batch of images fills the GPU with independent CTAs (5, 7 or 8 per SM,
// SegCfg below: while some CTAs sit in a barrier or a sequential reduction chain the others stream).
//
// PARITY MODE: the arithmetic follows the reference's compiled Eigen code (SEG.cpp =
// Segmentation/Segmentation/cython/src/LPboxADMMsolver.cpp): no FMA, row-sequential SpMV in ascending column order,
// reductions in Eigen's SSE2 order.  A reduction over n elements is four dependent chains of n/4 adds; the staging warps
// stream the products into a double-buffered shared-memory ring while ONE warp walks the chains (up to 7 reductions side by
// side, 4 lanes each).  What makes it fast: the matrix is read through an 8-entry row image fetched one element
// ahead (SegRowRef), element-wise passes are fused with the reductions that consume them (seg_fused_pass), and the streamed
// operands of the next element are prefetched into L2 (seg_pf).
//
// The CG matrix `temp_mat = 2A + (rho1+rho2) I` (SEG.cpp:784-786) is never materialised: off-diagonal entries are
// 2*a_ij (exact), the diagonal lives in `md` and is patched additively like the reference (SEG.cpp:1240-1243).
#pragma once
#include "lp_kernels.cuh"

namespace lpb {

struct SegInst {
    int n0, nnz0;            // capacities
    int n, nnz;              // current
    int iter, status, done, last_ret;
    int rhoUpdated;
    int n_ret;
    int xit_rows, xit_cols;
    int cur;                 // which half of the double-buffered arrays (CSR, b) is current (early fixing ping-pongs)
    int pad;
    double rho1, rho2, prho1, prho2, gamma, ratio;
    double std_obj, cur_obj, best_bin_obj, cconst;
    long long obj_len;
    double obj_ring[16];
    long long cg_iters, admm_iters;
};

struct SegView {
    int B, hist_cap;
    const long long *off_n;    // [B+1] element offsets of n-vectors (n0 rounded up to 4)
    const long long *off_nnz;  // [B+1] element offsets of colidx / val
    const long long *off_hist;
    double *x, *y1, *y2, *z1, *z2, *md, *invd, *r, *p, *t, *w;   // n-vectors
    double *b[2];              // double-buffered (early fixing rewrites it)
    int *rowptr[2];            // (off_n + i) offset, n0 + 1 entries per instance (stride n0r + 4 keeps room)
    void *colidx[2];           // general: int32 column index, fp64 value;  compact: int16 (column - row), int8 value --
    void *val[2];              // see SegFmt (values of the graph builder are small integers, neighbours are < 2^15 rows away)
    SegInst *st;
    double *hist;              // [cc][n0]
    int *left_idx, *ret_idx;
    double *ret_val;
    const double *pow_tab;     // pow_tab[k] = std::pow(k, 0.5) from the host libm, k <= max n0
    double *powv;              // [B] std::pow(n, 0.5) of the CURRENT n, refreshed by the early-fix kernel
    uint4 *ell_c;              // [off_n] compact format only: the row image the ADMM kernel reads -- 8 x int16 column distances
    uint2 *ell_a;              // [off_n]   (SEG_ELL_PAD = no entry) and 8 x int8 values per row, built from the CSR arrays (SegRowRef)
    int use_ell;               // every row has <= 8 stored entries: seg_admm_kernel<., ., true> reads ell_c / ell_a instead of the CSR arrays
    int *kidx;                 // [off_n] scratch: new index of a kept variable
    int *cnt;                  // [off_n] scratch: kept entries per new row / misc
};

struct SegLaunch {
    int iter_start, iter_end;
    int l2f, skip_done;
    int n_work;
    int *counter;
};

constexpr int SEG_T = 256;      // threads of the set-up / early-fix kernels and of the wide ADMM variant
constexpr int SEG_RMAX = 7;

// Launch shapes of seg_admm_kernel (one chain warp + T/32 - 1 staging warps per CTA; 40-42 resident warps per SM each):
//   T = 256: 5 CTAs/SM (740 images resident on 148 SMs), 224 staging threads per image, 48 registers -- fastest per image;
//   T = 192: 7 CTAs/SM (1036 resident), 160 staging threads per image, 40 registers (42 warps do not fit at 48);
//   T = 160: 8 CTAs/SM (1184 resident), 128 staging threads per image, 48 registers.
// The host picks the shape that needs fewer waves (e.g. the 1024-image batch of configs[2]: one wave instead of 1.38).
// BUF = doubles of the double-buffered product ring; a fused pass with R reductions stages FCH<R> elements per chunk (a multiple
// of the staging thread count): few reductions -> long chunks -> more rows in flight per thread between two barriers.
// Staging elements a thread works on side by side.  Measured on 375x500 images: 1 beats 2, 3 and 4 for every shape (+5 %): the
// row prefetch already overlaps the loads of consecutive elements and a second element in flight costs spills.
#ifndef SEG_UNROLL_W
#define SEG_UNROLL_W 1
#endif
#ifndef SEG_UNROLL_N
#define SEG_UNROLL_N 1
#endif
template <int T> struct SegCfg {
    static constexpr int STG = T - 32;
    static constexpr int MINB = (T == 256) ? 5 : (T == 192) ? 7 : 8;
    static constexpr int UNROLL = (T == 256) ? SEG_UNROLL_W : SEG_UNROLL_N;
    static constexpr int BUF = (T == 256) ? 4480 : (T == 192) ? 3840 : 3328;
    static constexpr int CH = (BUF / (2 * SEG_RMAX) < 256 ? BUF / (2 * SEG_RMAX) : 256) & ~3;   // seg_block_redux: products per reduction per chunk
};
constexpr int SEG_BUF_DOUBLES = SegCfg<SEG_T>::BUF;

// Storage format of A.  COMPACT = 3 bytes per stored entry instead of 12: the column index as int16 distance from the row and the
// value as int8 (exact: both conversions are lossless for the graphs of the reference's builder and for any user matrix the host
// has checked), so the streamed matrix traffic of every SpMV drops by 4x while the fp64 arithmetic sees the same numbers.
template <bool CMP> struct SegFmt { using CI = int; using AV = double; };
template <> struct SegFmt<true> { using CI = short; using AV = signed char; };
template <bool CMP>
__device__ __forceinline__ int seg_col(const typename SegFmt<CMP>::CI *__restrict__ ci, int k, int i) { return CMP ? i + (int)ci[k] : (int)ci[k]; }

// y_i = ((0 + m_i1 v_j1) + m_i2 v_j2) + ...  row i of (DIAG ? 2A with the diagonal replaced by md : A)
template <bool DIAG, bool CMP>
__device__ __forceinline__ double seg_row_dot(const int *__restrict__ rp, const typename SegFmt<CMP>::CI *__restrict__ ci,
                                              const typename SegFmt<CMP>::AV *__restrict__ av, const double *__restrict__ md,
                                              const double *__restrict__ v, int i) {
    double acc = 0.0;
    const int e = rp[i + 1];
    for (int k = rp[i]; k < e; ++k) {
        const int c = seg_col<CMP>(ci, k, i);
        double m = (double)av[k];
        if (DIAG) m = (c == i) ? md[i] : dM(2.0, m);
        acc = dA(acc, dM(m, v[c]));
    }
    return acc;
}

// What a staging thread holds of one row of A while it works on it, loaded ONE ELEMENT AHEAD by the fused passes so that the
// dependent chain of a row product is just "operands" instead of "rowptr -> column offsets / values -> operands":
//   general / compact CSR: the row bounds [s, e);
//   row image (compact format, every row <= 8 entries -- the reference builder's graphs have <= 7): the whole row, one 16-byte
//   and one 8-byte load: 8 x int16 column distance (SEG_ELL_PAD marks "no entry", padding sits at the end) + 8 x int8 value.
constexpr int SEG_ELL_PAD = -32768;
template <bool ELL> struct SegRowRef { int s, e; };
template <> struct SegRowRef<true> { uint4 c; uint2 a; };
template <bool ELL> struct SegRowSrc {
    const int *__restrict__ rp;
    __device__ __forceinline__ SegRowRef<false> load(int i) const { SegRowRef<false> r; r.s = rp[i]; r.e = rp[i + 1]; return r; }
};
template <> struct SegRowSrc<true> {
    const uint4 *__restrict__ c; const uint2 *__restrict__ a;
    __device__ __forceinline__ SegRowRef<true> load(int i) const { SegRowRef<true> r; r.c = __ldg(c + i); r.a = __ldg(a + i); return r; }
};

// y_i = ((0 + m_i1 v_j1) + m_i2 v_j2) + ... with the row already referenced by `rw`; mdi = md[i] (DIAG only).
template <bool DIAG, bool CMP>
__device__ __forceinline__ double seg_row_dot_ref(const SegRowRef<false> &rw, const typename SegFmt<CMP>::CI *__restrict__ ci,
                                                  const typename SegFmt<CMP>::AV *__restrict__ av, double mdi,
                                                  const double *__restrict__ v, int i) {
    double acc = 0.0;
    for (int k = rw.s; k < rw.e; ++k) {
        const int c = seg_col<CMP>(ci, k, i);
        double m = (double)av[k];
        if (DIAG) m = (c == i) ? mdi : dM(2.0, m);
        acc = dA(acc, dM(m, v[c]));
    }
    return acc;
}
template <bool DIAG, bool CMP>
__device__ __forceinline__ double seg_row_dot_ref(const SegRowRef<true> &rw, const typename SegFmt<CMP>::CI *__restrict__,
                                                  const typename SegFmt<CMP>::AV *__restrict__, double mdi,
                                                  const double *__restrict__ v, int i) {
    const unsigned cw[4] = {rw.c.x, rw.c.y, rw.c.z, rw.c.w}, aw[2] = {rw.a.x, rw.a.y};
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int d = (int)(short)(cw[k >> 1] >> (16 * (k & 1)));
        if (d != SEG_ELL_PAD) {                                  // same entries in the same order as the CSR walk
            double m = (double)(int)(signed char)(aw[k >> 2] >> (8 * (k & 3)));
            if (DIAG) m = (d == 0) ? mdi : dM(2.0, m);
            acc = dA(acc, dM(m, v[i + d]));
        }
    }
    return acc;
}

// Two products with one walk over row i of A: ax = (A x)_i and aw = (A 1[x >= 0.5])_i, each accumulated in row order.
template <bool CMP>
__device__ __forceinline__ void seg_row_dot2_ref(const SegRowRef<false> &rw, const typename SegFmt<CMP>::CI *__restrict__ ci,
                                                 const typename SegFmt<CMP>::AV *__restrict__ av, const double *__restrict__ v, int i,
                                                 double &ax, double &aw) {
    double a0 = 0.0, a1 = 0.0;
    for (int k = rw.s; k < rw.e; ++k) {
        const double m = (double)av[k], xc = v[seg_col<CMP>(ci, k, i)];
        a0 = dA(a0, dM(m, xc));
        a1 = dA(a1, dM(m, (xc >= 0.5) ? 1.0 : 0.0));
    }
    ax = a0; aw = a1;
}
template <bool CMP>
__device__ __forceinline__ void seg_row_dot2_ref(const SegRowRef<true> &rw, const typename SegFmt<CMP>::CI *__restrict__,
                                                 const typename SegFmt<CMP>::AV *__restrict__, const double *__restrict__ v, int i,
                                                 double &ax, double &aw) {
    const unsigned cw[4] = {rw.c.x, rw.c.y, rw.c.z, rw.c.w}, aws[2] = {rw.a.x, rw.a.y};
    double a0 = 0.0, a1 = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int d = (int)(short)(cw[k >> 1] >> (16 * (k & 1)));
        if (d != SEG_ELL_PAD) {
            const double m = (double)(int)(signed char)(aws[k >> 2] >> (8 * (k & 3))), xc = v[i + d];
            a0 = dA(a0, dM(m, xc));
            a1 = dA(a1, dM(m, (xc >= 0.5) ? 1.0 : 0.0));
        }
    }
    ax = a0; aw = a1;
}

// Software prefetch of the streamed operands of the element a staging thread will work on next (into L2: the request holds no
// register, and the demand load one element later finds the line on chip instead of in HBM): +15 % at a full wave.  Measured and
// not kept: two or three elements ahead (same), also prefetching the farthest forward neighbour of the gathers (-1 %).
#ifndef SEG_PF_DIST
#define SEG_PF_DIST 1
#endif
__device__ __forceinline__ void seg_pf(const void *p) {
#ifndef SEG_NO_PREFETCH
    asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
#endif
}

// Walks one Eigen chain: adds src[i], src[i+4], ... (indices < lim) to acc IN ORDER.  The next four terms are fetched into
// loop-carried registers before the four dependent adds; otherwise ptxas, short of registers (5 CTAs/SM), serialises
// load -> add -> load on one register and every step pays the shared-memory latency on top of the fp64 add latency.
__device__ __forceinline__ double seg_chain(const double *src, int i, int lim, double acc) {
    int left = lim > i ? (lim - i + 3) >> 2 : 0;
    const double *pv = src + i;
    if (left >= 4) {
        double t0 = pv[0], t1 = pv[4], t2 = pv[8], t3 = pv[12];
        pv += 16; left -= 4;
#pragma unroll 1
        while (left >= 4) {
            const double u0 = pv[0], u1 = pv[4], u2 = pv[8], u3 = pv[12];
            acc = dA(acc, t0); acc = dA(acc, t1); acc = dA(acc, t2); acc = dA(acc, t3);
            t0 = u0; t1 = u1; t2 = u2; t3 = u3;
            pv += 16; left -= 4;
        }
        acc = dA(acc, t0); acc = dA(acc, t1); acc = dA(acc, t2); acc = dA(acc, t3);
    }
    for (; left > 0; --left, pv += 4) acc = dA(acc, pv[0]);
    return acc;
}

// Block-cooperative Eigen-order reduction of R product streams prod(q, i), i < n.  Results in sc[0..R).
// buf: shared, 2 * R * SegCfg<T>::CH doubles.  All T threads must call.
template <int R, int T = SEG_T, typename F>
__device__ __forceinline__ void seg_block_redux(F prod, int n, double *buf, double *sc) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int RW = T / 32 - 1, CH = SegCfg<T>::CH;
    const int a2 = n & ~3, a1 = n & ~1;
    const int q = lane >> 2, k = lane & 3;
    const int nch = (a2 + CH - 1) / CH;
    auto stage = [&](int c) {
        double *dst = buf + (size_t)(c & 1) * R * CH;
        const int base = c * CH;
        const int lim = min(CH, a2 - base);
        for (int idx = tid; idx < R * CH; idx += T - 32) {          // every warp but the reduction warp
            const int qq = idx / CH, i = idx - qq * CH;
            if (i < lim) dst[qq * CH + i] = prod(qq, base + i);
        }
    };
    double acc = 0.0;
    __syncthreads();   // every reader of the previous results in sc[] is done; operands written by other threads are visible
    if (a1 > 2) {
        if (warp != RW) stage(0);
        __syncthreads();
        for (int c = 0; c < nch; ++c) {
            if (warp != RW) { if (c + 1 < nch) stage(c + 1); }
            else if (q < R) {
                const double *src = buf + (size_t)(c & 1) * R * CH + q * CH;
                const int lim = min(CH, a2 - c * CH);
                int i = k;
                if (c == 0) { acc = src[k]; i = 4 + k; }
                acc = seg_chain(src, i, lim, acc);
            }
            __syncthreads();
        }
    }
    if (warp == RW) {
        const int qq = q < R ? q : 0;
        double res;
        if (a1 > 2) {
            double hi = __shfl_down_sync(0xffffffffu, acc, 2);
            double l = dA(acc, hi);
            if (a1 > a2 && k < 2) l = dA(l, prod(qq, a2 + k));
            double l1 = __shfl_down_sync(0xffffffffu, l, 1);
            res = dA(l, l1);
        } else if (a1 == 2) {
            res = dA(prod(qq, 0), prod(qq, 1));
        } else {
            res = (n > 0) ? prod(qq, 0) : 0.0;
        }
        if ((n & 1) && n > 1) res = dA(res, prod(qq, n - 1));
        if (k == 0 && q < R) sc[q] = res;
    }
    __syncthreads();
}

// Fused streaming pass + Eigen-order reductions.  body(i, rw, v) is called EXACTLY ONCE for every i < n (in chunk order) by
// the staging warps: it performs the element's work (global loads / stores) and returns the R products of element i; with
// ROWS, rw references row i of A (SegRowRef, loaded one element ahead).  The products go to a double-buffered shared-memory
// ring; the reduction warp walks the four chains of each reduction one chunk behind the producers, so the streaming work and
// the sequential chains overlap.  Results in sc[0..R).  The chunk length only decides how the work is staged, never the order
// of the additions.
template <int T, int R, bool ROWS, bool ELL, typename Body, typename Pre>
__device__ __forceinline__ void seg_fused_pass(Body body, Pre pre, const SegRowSrc<ELL> &rows, int n, double *buf, double *sc) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int RW = T / 32 - 1, STG = T - 32;
    constexpr int FCH = (SegCfg<T>::BUF / (2 * R)) / STG * STG;
    static_assert(FCH >= STG, "product ring too small for one element per staging thread");
    // short vectors: about 8 chunks (so that staging and chain overlap), at least one element per staging thread
    const int fch = (n >= 8 * FCH) ? FCH : min(FCH, max(STG, (n / 8) / STG * STG));
    const int a2 = n & ~3, a1 = n & ~1;
    const int q = lane >> 2, k = lane & 3;
    const int nch = (n + fch - 1) / fch;
    // A staging thread works on the elements tid, tid + STG, tid + 2 STG, ... (fch is a multiple of STG, so the sequence runs on
    // across chunk boundaries); the row of the NEXT element is requested before the current one is processed.
    SegRowRef<ELL> nx{};
    if (ROWS && warp != RW && tid < n) nx = rows.load(tid);
    auto stage = [&](int c) {
        double *dst = buf + (size_t)(c & 1) * R * fch;
        const int base = c * fch;
        const int lim = min(fch, n - base);
#pragma unroll (SegCfg<T>::UNROLL)
        for (int idx = tid; idx < lim; idx += STG) {
            const int i = base + idx;
            const SegRowRef<ELL> rw = nx;
            if (ROWS && i + STG < n) nx = rows.load(i + STG);
            if (i + SEG_PF_DIST * STG < n) pre(i + SEG_PF_DIST * STG);
            double v[R];
            body(i, rw, v);
#pragma unroll
            for (int r = 0; r < R; ++r) dst[r * fch + idx] = v[r];
        }
    };
    double acc = 0.0;
    __syncthreads();   // previous readers of sc[] are done; operands written by other threads are visible
    if (warp != RW) stage(0);
    __syncthreads();
    for (int c = 0; c < nch; ++c) {
        if (warp != RW) { if (c + 1 < nch) stage(c + 1); }
        else if (q < R && a1 > 2) {
            const double *src = buf + (size_t)(c & 1) * R * fch + q * fch;
            const int lim = min(fch, a2 - c * fch);             // chain part only (elements < a2)
            int i = k;
            if (c == 0) { acc = src[k]; i = 4 + k; }
            acc = seg_chain(src, i, lim, acc);
        }
        __syncthreads();
    }
    if (warp == RW) {
        // products of the tail elements (index >= a2) sit in the ring slot of the chunk that contains them
        const int qq = q < R ? q : 0;
        auto at = [&](int i) { const int c = i / fch; return buf[(size_t)(c & 1) * R * fch + qq * fch + (i - c * fch)]; };
        // only the last two chunks are still resident: all indices >= a2 - and, for n < 4, indices 0..n-1 - are in them
        double res;
        if (a1 > 2) {
            double hi = __shfl_down_sync(0xffffffffu, acc, 2);
            double l = dA(acc, hi);
            if (a1 > a2 && k < 2) l = dA(l, at(a2 + k));
            double l1 = __shfl_down_sync(0xffffffffu, l, 1);
            res = dA(l, l1);
        } else if (a1 == 2) {
            res = dA(at(0), at(1));
        } else {
            res = (n > 0) ? at(0) : 0.0;
        }
        if ((n & 1) && n > 1) res = dA(res, at(n - 1));
        if (k == 0 && q < R) sc[q] = res;
    }
    __syncthreads();
}

template <bool CMP, int T, bool ELL>
__global__ void __launch_bounds__(T, SegCfg<T>::MINB)
seg_admm_kernel(SegView sv, Params pr, SegLaunch la) {
    static_assert(CMP || !ELL, "the row image exists for the compact format only");
    using Row = SegRowRef<ELL>;
    using CI = typename SegFmt<CMP>::CI;
    using AV = typename SegFmt<CMP>::AV;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *buf = reinterpret_cast<double *>(smem_raw);                 // product ring, SegCfg<T>::BUF doubles
    double *sc = buf + SegCfg<T>::BUF;                                 // [8]
    double *ring = sc + 8;                                              // [16]
    __shared__ int s_work;
    const int tid = threadIdx.x;
    for (;;) {
        if (tid == 0) s_work = atomicAdd(la.counter, 1);
        __syncthreads();
        const int wk = s_work;
        __syncthreads();
        if (wk >= la.n_work) break;
        const int inst = wk;
        SegInst *st = sv.st + inst;
        if ((la.skip_done && st->done) || st->n == 0) continue;
        const int n = st->n, cur = st->cur;
        const long long on = sv.off_n[inst], oz = sv.off_nnz[inst];
        double *x = sv.x + on, *y1 = sv.y1 + on, *y2 = sv.y2 + on, *z1 = sv.z1 + on, *z2 = sv.z2 + on, *md = sv.md + on,
               *invd = sv.invd + on, *r = sv.r + on, *p = sv.p + on, *t = sv.t + on, *w = sv.w + on;
        const double *__restrict__ b = sv.b[cur] + on;
        const int *__restrict__ rp = sv.rowptr[cur] + on + 4 * inst;      // n0r + 4 ints reserved per instance
        const CI *__restrict__ ci = reinterpret_cast<const CI *>(sv.colidx[cur]) + oz;
        const AV *__restrict__ av = reinterpret_cast<const AV *>(sv.val[cur]) + oz;
        SegRowSrc<ELL> rows;
        if constexpr (ELL) { rows.c = sv.ell_c + on; rows.a = sv.ell_a + on; } else { rows.rp = rp; }
        double rho1 = st->rho1, rho2 = st->rho2, prho1 = st->prho1, prho2 = st->prho2, gamma = st->gamma, ratio = st->ratio,
               std_obj = st->std_obj, cur_obj = st->cur_obj, best_bin_obj = st->best_bin_obj;
        int rhoUpdated = st->rhoUpdated;
        long long obj_len = st->obj_len, cg_total = 0, admm_total = 0;
        if (tid < 16) ring[tid] = st->obj_ring[tid];
        const double pow_n = sv.powv[inst];
        __syncthreads();

        int status = RUNNING, iter = la.iter_start, cc = 0;
        for (; iter < la.iter_end; ++iter) {
            // ---- pass 1: y1, y2 pre-image (SEG.cpp:1223-1234) + ||y||^2 ---------------------------------------------
            seg_fused_pass<T, 1, false, ELL>([&](int i, const Row &, double (&v)[1]) {
                const double xi = x[i];
                double tt = dA(xi, dD(z1[i], rho1));
                y1[i] = (tt > 1.0) ? 1.0 : ((tt < 0.0) ? 0.0 : tt);
                const double sh = dS(dA(xi, dD(z2[i], rho2)), 0.5);
                y2[i] = sh;
                v[0] = dM(sh, sh);
            }, [&](int j) { seg_pf(x + j); seg_pf(z1 + j); seg_pf(z2 + j); }, rows, n, buf, sc);
            const double den = dM(2.0, sqrt(sc[0]));
            // ---- pass 2+3: diagonal patch (:1240-1243), preconditioner (:1252-1255), y2, rhs (:1246), warm start x = y1, and the
            // PCG prologue (SEG.cpp:272-342) r = rhs - M x, p = invd r with rhs.rhs, r.r, r.p.  One pass: every quantity of element i
            // but the row product depends on element i only, and the row product gathers x = y1, which pass 1 has completed.
            const bool patch = (iter != 0 && rhoUpdated), refresh = rhoUpdated != 0;
            const double dpatch = dM(dA(prho1, prho2), ratio);
            seg_fused_pass<T, 3, true, ELL>([&](int i, const Row &rw, double (&v)[3]) {
                double mdi = md[i];
                if (patch) { mdi = dA(mdi, dpatch); md[i] = mdi; }
                double idi;
                if (refresh) { idi = (mdi != 0.0) ? dD(1.0, mdi) : 1.0; invd[i] = idi; } else idi = invd[i];
                const double y1i = y1[i];
                const double y2v = dA(dD(dM(y2[i], pow_n), den), 0.5);
                y2[i] = y2v;
                const double rhs = dS(dA(dM(rho1, y1i), dM(rho2, y2v)), dA(dA(b[i], z1[i]), z2[i]));
                const double rr = dS(rhs, seg_row_dot_ref<true, CMP>(rw, ci, av, mdi, y1, i));
                const double pp = dM(idi, rr);
                x[i] = y1i; r[i] = rr; p[i] = pp;
                v[0] = dM(rhs, rhs); v[1] = dM(rr, rr); v[2] = dM(rr, pp);
            }, [&](int j) { seg_pf(md + j); seg_pf(invd + j); seg_pf(y1 + j); seg_pf(y2 + j); seg_pf(b + j); seg_pf(z1 + j); seg_pf(z2 + j); }, rows, n, buf, sc);
            rhoUpdated = 0;
            const double rhsNorm2 = sc[0];
            int cg_it = 0;
            if (rhsNorm2 == 0.0) {
                for (int i = tid; i < n; i += T) x[i] = 0.0;
            } else {
                double threshold = dM(dM(pr.pcg_tol, pr.pcg_tol), rhsNorm2);
                if (!(threshold > DBL_MIN)) threshold = DBL_MIN;
                double r2 = sc[1], absNew = sc[2];
                if (!(r2 < threshold)) {
                    while (cg_it < pr.pcg_maxiters) {
                        // tmp = M p fused with p.dot(tmp)
                        seg_fused_pass<T, 1, true, ELL>([&](int i, const Row &rw, double (&v)[1]) {
                            const double ti = seg_row_dot_ref<true, CMP>(rw, ci, av, md[i], p, i);
                            t[i] = ti;
                            v[0] = dM(p[i], ti);
                        }, [&](int j) { seg_pf(md + j); seg_pf(p + j); }, rows, n, buf, sc);
                        const double alpha = dD(absNew, sc[0]);
                        // x += alpha p; r -= alpha tmp; z = invd r fused with r.r and r.z
                        seg_fused_pass<T, 2, false, ELL>([&](int i, const Row &, double (&v)[2]) {
                            x[i] = dA(x[i], dM(alpha, p[i]));
                            const double rr = dS(r[i], dM(alpha, t[i]));
                            const double zz = dM(invd[i], rr);
                            r[i] = rr; t[i] = zz;
                            v[0] = dM(rr, rr); v[1] = dM(rr, zz);
                        }, [&](int j) { seg_pf(x + j); seg_pf(p + j); seg_pf(r + j); seg_pf(t + j); seg_pf(invd + j); }, rows, n, buf, sc);
                        r2 = sc[0];
                        if (r2 < threshold) { cg_it++; break; }
                        const double absOld = absNew;
                        absNew = sc[1];
                        const double beta = dD(absNew, absOld);
#pragma unroll 4
                        for (int i = tid; i < n; i += T) p[i] = dA(t[i], dM(beta, p[i]));
                        cg_it++;
                    }
                }
            }
            cg_total += cg_it; admm_total += 1;
            // ---- pass: history (SEG.cpp:1131-1134), duals (:1280-1281), indicator, A x; x.x, (x-y1)^2, (x-y2)^2, x.Ax, b.x ----
            double *h = nullptr;
            if (la.l2f && sv.hist_cap > 0) { if (cc < sv.hist_cap) h = sv.hist + sv.off_hist[inst] + (long long)cc * st->n0; cc++; }
            {
                const double g1 = dM(gamma, rho1), g2 = dM(gamma, rho2);
                // one walk over row i yields (A x)_i and (A 1[x >= 0.5])_i: the second product (SEG.cpp:1323-1326) needs no pass of its own
                seg_fused_pass<T, 7, true, ELL>([&](int i, const Row &rw, double (&v)[7]) {
                    const double xi = x[i];
                    if (h) h[i] = xi;
                    const double d1 = dS(xi, y1[i]), d2 = dS(xi, y2[i]);
                    z1[i] = dA(z1[i], dM(g1, d1));
                    z2[i] = dA(z2[i], dM(g2, d2));
                    const double wi = (xi >= 0.5) ? 1.0 : 0.0, bi = b[i];
                    double ax, aw;
                    seg_row_dot2_ref<CMP>(rw, ci, av, x, i, ax, aw);
                    v[0] = dM(xi, xi); v[1] = dM(d1, d1); v[2] = dM(d2, d2); v[3] = dM(xi, ax); v[4] = dM(bi, xi);
                    v[5] = dM(wi, aw); v[6] = dM(bi, wi);
                }, [&](int j) { seg_pf(x + j); seg_pf(y1 + j); seg_pf(y2 + j); seg_pf(z1 + j); seg_pf(z2 + j); seg_pf(b + j); }, rows, n, buf, sc);
            }
            const double nx2 = sc[0], d12 = sc[1], d22 = sc[2], obj_val = dA(sc[3], sc[4]);   // compute_cost: val + val2
            const double bin_val = dA(sc[5], sc[6]);                                         // idx.A idx + b.idx
            {
                double temp0 = sqrt(nx2);
                if (!(temp0 > 2.2204e-16)) temp0 = 2.2204e-16;
                const double c1 = dD(sqrt(d12), temp0), c2 = dD(sqrt(d22), temp0);
                if (c1 <= pr.stop_threshold && c2 <= pr.stop_threshold) { status = STOP_Y; break; }   // SEG.cpp:1288-1292
            }
            if ((iter + 1) % pr.rho_change_step == 0) {                  // :1295-1303
                prho1 = rho1; prho2 = rho2;
                rho1 = dM(pr.learning_fact, rho1); rho2 = dM(pr.learning_fact, rho2);
                double g = dM(gamma, pr.gamma_factor);
                gamma = (g < 1.0) ? 1.0 : g;
                rhoUpdated = 1;
                ratio = dS(pr.learning_fact, 1.0);
            }
            {
                const double obj = obj_val;
                double so = std_obj;
                if (obj_len + 1 >= (long long)pr.history_size) so = std_obj_after_push(ring, obj_len, obj, pr.history_size);
                __syncthreads();
                if (tid == 0) ring[obj_len & 15] = obj;
                obj_len++;
                std_obj = so;
                __syncthreads();
                if (std_obj <= pr.std_threshold) { status = STOP_STD; break; }   // :1313-1319
            }
            cur_obj = bin_val;                                           // :1323-1326
            if (best_bin_obj >= cur_obj) best_bin_obj = cur_obj;
        }
        __syncthreads();
        if (!la.l2f) {
            // legacy epilogue (SEG.cpp:1366-1367): cur_obj = compute_cost(1[x >= 0.5])
            for (int i = tid; i < n; i += T) w[i] = (x[i] >= 0.5) ? 1.0 : 0.0;
            __syncthreads();
            for (int i = tid; i < n; i += T) r[i] = seg_row_dot<false, CMP>(rp, ci, av, md, w, i);
            __syncthreads();
            seg_block_redux<2, T>([&](int q, int i) { return dM(q == 0 ? w[i] : b[i], q == 0 ? r[i] : w[i]); }, n, buf, sc);
            cur_obj = dA(sc[0], sc[1]);
        }
        if (tid < 16) st->obj_ring[tid] = ring[tid];
        if (tid == 0) {
            st->rho1 = rho1; st->rho2 = rho2; st->prho1 = prho1; st->prho2 = prho2; st->gamma = gamma; st->ratio = ratio;
            st->std_obj = std_obj; st->cur_obj = cur_obj; st->best_bin_obj = best_bin_obj; st->rhoUpdated = rhoUpdated;
            st->obj_len = obj_len; st->cg_iters += cg_total; st->admm_iters += admm_total; st->iter = iter; st->status = status;
            const int ret = la.l2f ? (status != RUNNING ? 1 : 0) : 0;
            st->last_ret = ret;
            st->done = (status != RUNNING) ? 1 : 0;
            if (la.l2f) st->xit_cols = cc;
        }
        __syncthreads();
    }
}

// ADMM_bqp_unconstrained_init (SEG.cpp:747-810) for every image: x = x0 (zeros), y = x, z = 0, md = 2 a_ii + (rho1+rho2),
// best_bin_obj = compute_cost(x0).  One CTA per image.
template <bool CMP>
__global__ void __launch_bounds__(SEG_T)
seg_setup_kernel(SegView sv, Params pr, int use_x0) {
    using CI = typename SegFmt<CMP>::CI;
    using AV = typename SegFmt<CMP>::AV;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *buf = reinterpret_cast<double *>(smem_raw);
    double *sc = buf + SEG_BUF_DOUBLES;
    const int inst = blockIdx.x, tid = threadIdx.x;
    SegInst *st = sv.st + inst;
    const int n = st->n, cur = st->cur;
    const long long on = sv.off_n[inst], oz = sv.off_nnz[inst];
    const int *rp = sv.rowptr[cur] + on + 4 * inst;
    const CI *ci = reinterpret_cast<const CI *>(sv.colidx[cur]) + oz;
    const AV *av = reinterpret_cast<const AV *>(sv.val[cur]) + oz;
    const double *b = sv.b[cur] + on;
    double *x = sv.x + on, *t = sv.t + on;
    const double rho = pr.initial_rho;
    for (int i = tid; i < n; i += SEG_T) {
        const double x0 = use_x0 ? x[i] : 0.0;
        x[i] = x0; sv.y1[on + i] = x0; sv.y2[on + i] = x0; sv.z1[on + i] = 0.0; sv.z2[on + i] = 0.0;
        sv.left_idx[on + i] = i;
        double d = 0.0;
        for (int k = rp[i]; k < rp[i + 1]; ++k) if (seg_col<CMP>(ci, k, i) == i) d = dM(2.0, (double)av[k]);
        sv.md[on + i] = dA(d, dA(rho, rho));                             // temp_mat = 2A; diag += rho1 + rho2
    }
    __syncthreads();
    for (int i = tid; i < n; i += SEG_T) t[i] = seg_row_dot<false, CMP>(rp, ci, av, nullptr, x, i);
    __syncthreads();
    seg_block_redux<2>([&](int q, int i) { return dM(q == 0 ? x[i] : b[i], q == 0 ? t[i] : x[i]); }, n, buf, sc);
    if (tid == 0) {
        st->rho1 = st->rho2 = st->prho1 = st->prho2 = rho; st->gamma = pr.gamma_val; st->ratio = 0.0; st->rhoUpdated = 1;
        st->std_obj = 1.0; st->cur_obj = 0.0; st->best_bin_obj = dA(sc[0], sc[1]); st->obj_len = 0; st->cg_iters = 0; st->admm_iters = 0;
        st->iter = 0; st->status = RUNNING; st->done = 0; st->last_ret = 0; st->n_ret = 0; st->xit_rows = 0; st->xit_cols = 0;
        for (int k = 0; k < 16; ++k) st->obj_ring[k] = 0.0;
    }
}

// Row image of the compact format (SegRowRef<true>) from the current CSR arrays: one thread per row.  *flag is set when a row
// has more than 8 stored entries or a distance equal to the padding marker (the batch then keeps using the CSR arrays).
__global__ void __launch_bounds__(SEG_T)
seg_ell_build_kernel(SegView sv, int skip_done, int *flag) {
    const int inst = blockIdx.x;                         // grid = (images, row chunks): the x dimension holds any batch size
    const SegInst *st = sv.st + inst;
    if (skip_done && st->done) return;
    const int n = st->n, cur = st->cur;
    const long long on = sv.off_n[inst], oz = sv.off_nnz[inst];
    const int *__restrict__ rp = sv.rowptr[cur] + on + 4 * inst;
    const short *__restrict__ ci = reinterpret_cast<const short *>(sv.colidx[cur]) + oz;
    const signed char *__restrict__ av = reinterpret_cast<const signed char *>(sv.val[cur]) + oz;
    for (int i = blockIdx.y * SEG_T + threadIdx.x; i < n; i += gridDim.y * SEG_T) {
        const int s = rp[i];
        int len = rp[i + 1] - s;
        if (len > 8) { *flag = 1; len = 8; }
        unsigned cw[4] = {0, 0, 0, 0}, aw[2] = {0, 0};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int d = SEG_ELL_PAD, a = 0;
            if (k < len) { d = ci[s + k]; a = av[s + k]; if (d == SEG_ELL_PAD) *flag = 1; }
            cw[k >> 1] |= (unsigned)(d & 0xffff) << (16 * (k & 1));
            aw[k >> 2] |= (unsigned)(a & 0xff) << (8 * (k & 3));
        }
        sv.ell_c[on + i] = make_uint4(cw[0], cw[1], cw[2], cw[3]);
        sv.ell_a[on + i] = make_uint2(aw[0], aw[1]);
    }
}


// ---- device-side early fixing for the unconstrained form: ADMM_bqp_unconstrained_l2f prologue (SEG.cpp:932-1090) -------
// One CTA per image.  A <- A[keep,keep] (Ma), b <- 2 A[keep,fix] x_fix + b[keep] (:1044-1052), x,y1,y2,z1,z2 gathered,
// left_idx / ret_idx / ret_val bookkeeping, temp_mat diagonal rebuilt at the CURRENT rho (:1054-1057).  The compressed
// arrays and b ping-pong between two buffers (no in-place hazards); n-vectors are gathered through the scratch vector r.
__device__ __forceinline__ int seg_block_exscan_global(int *data, int n, int *s_part) {
    // in-place exclusive scan of a global int array by the whole CTA; returns the total
    const int tid = threadIdx.x;
    constexpr int ITEMS = 8;
    int carry = 0;
    for (int base = 0; base < n; base += SEG_T * ITEMS) {
        int v[ITEMS], sum = 0;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) { const int i = base + tid * ITEMS + k; v[k] = (i < n) ? data[i] : 0; sum += v[k]; }
        s_part[tid] = sum;
        __syncthreads();
        if (tid == 0) { int run = 0; for (int t2 = 0; t2 < SEG_T; ++t2) { const int q = s_part[t2]; s_part[t2] = run; run += q; } s_part[SEG_T] = run; }
        __syncthreads();
        int run = carry + s_part[tid];
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) { const int i = base + tid * ITEMS + k; if (i < n) data[i] = run; run += v[k]; }
        carry += s_part[SEG_T];
        __syncthreads();
    }
    return carry;
}

template <bool CMP>
__global__ void __launch_bounds__(SEG_T)
seg_fix_kernel(SegView sv, Params pr, const double *__restrict__ vec, const long long *__restrict__ off_vec, const int *__restrict__ num,
               int skip_done) {
    __shared__ int s_part[SEG_T + 1];
    const int inst = blockIdx.x, tid = threadIdx.x;
    SegInst *st = sv.st + inst;
    if (skip_done && st->done) return;
    const int n = st->n;
    const int fn = num ? num[inst] : 0;
    if (fn == 0 || n == 0) { if (tid == 0) { st->xit_rows = n; st->xit_cols = 0; } return; }
    const int cur = st->cur, nxt = cur ^ 1;
    const long long on = sv.off_n[inst], oz = sv.off_nnz[inst];
    const double *v = vec + off_vec[inst];
    using CI = typename SegFmt<CMP>::CI;
    using AV = typename SegFmt<CMP>::AV;
    const int *rp = sv.rowptr[cur] + on + 4 * inst;
    const CI *ci = reinterpret_cast<const CI *>(sv.colidx[cur]) + oz;
    const AV *av = reinterpret_cast<const AV *>(sv.val[cur]) + oz;
    const double *b = sv.b[cur] + on;
    int *rp2 = sv.rowptr[nxt] + on + 4 * inst;
    CI *ci2 = reinterpret_cast<CI *>(sv.colidx[nxt]) + oz;
    AV *av2 = reinterpret_cast<AV *>(sv.val[nxt]) + oz;
    double *b2 = sv.b[nxt] + on;
    int *kidx = sv.kidx + on, *cnt = sv.cnt + on;
    auto is_fixed = [&](int i) { const double t = v[i]; return t == 1.0 || t == 0.0; };
    for (int i = tid; i < n; i += SEG_T) kidx[i] = is_fixed(i) ? 0 : 1;
    __syncthreads();
    const int k_tot = seg_block_exscan_global(kidx, n, s_part);
    const int j_tot = n - k_tot, n_ret = st->n_ret;
    // bookkeeping (SEG.cpp:1017-1026): left_idx compacted through cnt as scratch
    for (int i = tid; i < n; i += SEG_T) cnt[i] = sv.left_idx[on + i];
    __syncthreads();
    for (int i = tid; i < n; i += SEG_T) {
        if (is_fixed(i)) { const int q = n_ret + (i - kidx[i]); sv.ret_idx[on + q] = cnt[i]; sv.ret_val[on + q] = v[i]; }
        else sv.left_idx[on + kidx[i]] = cnt[i];
    }
    __syncthreads();
    if (k_tot == 0) {                                                         // :1028-1032
        if (tid == 0) { st->n = 0; st->nnz = 0; st->n_ret = n_ret + j_tot; st->status = STOP_EMPTY; st->last_ret = 1; st->done = 1; st->xit_rows = 0; st->xit_cols = 0; }
        return;
    }
    // gathers (:1035-1041) through the scratch vector r
    double *vecs[5] = {sv.x + on, sv.y1 + on, sv.y2 + on, sv.z1 + on, sv.z2 + on};
    double *scr = sv.r + on;
    for (int a = 0; a < 5; ++a) {
        for (int i = tid; i < n; i += SEG_T) scr[i] = vecs[a][i];
        __syncthreads();
        for (int i = tid; i < n; i += SEG_T) if (!is_fixed(i)) vecs[a][kidx[i]] = scr[i];
        __syncthreads();
    }
    // Ma = A[keep,keep], b = 2 (Mb x2) + b1 (:973-1015, :1044-1052); md = 2 a_ii + (rho1 + rho2) (:1054-1057)
    for (int i = tid; i < n; i += SEG_T) cnt[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += SEG_T) {
        if (is_fixed(i)) continue;
        int c = 0;
        for (int k = rp[i]; k < rp[i + 1]; ++k) c += is_fixed(seg_col<CMP>(ci, k, i)) ? 0 : 1;
        cnt[kidx[i]] = c;
    }
    __syncthreads();
    const int nnz_new = seg_block_exscan_global(cnt, k_tot, s_part);
    const double rr = dA(st->rho1, st->rho2);
    for (int i = tid; i < n; i += SEG_T) {
        if (is_fixed(i)) continue;
        const int r = kidx[i];
        int q = cnt[r];
        rp2[r] = q;
        double acc = 0.0, dg = 0.0;
        for (int k = rp[i]; k < rp[i + 1]; ++k) {
            const int c = seg_col<CMP>(ci, k, i);
            const double a = (double)av[k];
            if (is_fixed(c)) acc = dA(acc, dM(a, v[c]));
            else { ci2[q] = (CI)(CMP ? kidx[c] - r : kidx[c]); av2[q] = av[k]; q++; if (c == i) dg = dM(2.0, a); }   // kept neighbours only move closer
        }
        b2[r] = dA(dM(2.0, acc), b[i]);
        sv.md[on + r] = dA(dg, rr);
    }
    if (tid == 0) {
        rp2[k_tot] = nnz_new;
        st->n = k_tot; st->nnz = nnz_new; st->n_ret = n_ret + j_tot; st->cur = nxt; st->xit_rows = k_tot; st->xit_cols = 0;
        sv.powv[inst] = sv.pow_tab[k_tot];
    }
}

}  // namespace lpb
