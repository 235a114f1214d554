// Device-side glue between the window kernel and the early-fixing policy (LP.trainer:510-535 without host copies):
//   * lp_policy_input_kernel: iterate history [iteration][n0] (fp64) -> packed policy input [row][ws] (fp32), i.e. the
//     reference's `xiters.reshape(n_left, 20, ws/20).astype(float32)` (LP.trainer:524-530) for all active instances;
//   * lp_threshold_kernel: deter_fix_2 (LP.trainer:101-135) + the "n <= 10 -> no fix" rule (:533-535) on the scores.
#pragma once
#include "lp_kernels.cuh"

namespace lpb {

// Device-side active list of the window loop (LP.trainer:510-535 without a host scan): one CTA compacts the instances that
// are still running (not done, n > 0) into `active`, their row offsets (exclusive scan of n) into `row_off`, and writes
// meta = {n_active, total rows, largest n}.  The kernels below take n_active from meta and exit early beyond it, so the host
// launches them for the worst case and only reads back the 24 bytes of meta per window (to size the policy launch).
struct L2fMeta { long long n_active, rows, max_n; };
__global__ void __launch_bounds__(1024) lp_active_scan_kernel(BatchView bv, int *__restrict__ active, long long *__restrict__ row_off,
                                                              L2fMeta *__restrict__ meta) {
    __shared__ long long s_rows[1024];
    __shared__ int s_cnt[1024];
    __shared__ int s_max;
    const int tid = threadIdx.x, B = bv.B;
    const int per = (B + 1023) / 1024;
    const int beg = min(tid * per, B), end = min(beg + per, B);
    int cnt = 0, mx = 0;
    long long rows = 0;
    for (int i = beg; i < end; ++i) {
        const InstState &st = bv.st[i];
        if (!st.done && st.n != 0) { cnt++; rows += st.n; mx = max(mx, st.n); }
    }
    s_cnt[tid] = cnt; s_rows[tid] = rows;
    if (tid == 0) s_max = 0;
    __syncthreads();
    atomicMax(&s_max, mx);
    if (tid == 0) {                                   // 1024 partials: a serial scan is cheap next to a window of ADMM iterations
        int c = 0; long long r = 0;
        for (int t = 0; t < 1024; ++t) { const int ct = s_cnt[t]; const long long rt = s_rows[t]; s_cnt[t] = c; s_rows[t] = r; c += ct; r += rt; }
        meta->n_active = c; meta->rows = r;
        row_off[c] = r;
    }
    __syncthreads();
    if (tid == 0) meta->max_n = s_max;
    int c = s_cnt[tid];
    long long r = s_rows[tid];
    for (int i = beg; i < end; ++i) {
        const InstState &st = bv.st[i];
        if (!st.done && st.n != 0) { active[c] = i; row_off[c] = r; c++; r += st.n; }
    }
}

// grid = (ceil(max_rows/32), n_active); block = (32, 8).  Tile transpose through shared memory so that both the reads
// (along the variable index) and the writes (along the iteration index) are coalesced.
__global__ void lp_policy_input_kernel(BatchView bv, const int *__restrict__ active, const long long *__restrict__ row_off,
                                       int ws, float *__restrict__ out, const L2fMeta *__restrict__ meta) {
    __shared__ float tile[32][33];
    if (meta && (long long)blockIdx.y >= meta->n_active) return;       // launched for the worst case (device-side active list)
    const int inst = active[blockIdx.y];
    const InstState *st = bv.st + inst;
    const int n = st->n, n0 = st->n0, cols = min(st->xit_cols, min(ws, bv.hist_cap));
    const int r0 = blockIdx.x * 32;
    if (r0 >= n) return;
    const double *h = bv.hist + bv.off_hist[inst];
    float *o = out + row_off[blockIdx.y] * ws;
    for (int c0 = 0; c0 < ws; c0 += 32) {
        for (int cy = threadIdx.y; cy < 32; cy += 8) {
            const int c = c0 + cy, r = r0 + threadIdx.x;
            tile[cy][threadIdx.x] = (c < cols && r < n) ? (float)h[(long long)c * n0 + r] : 0.0f;   // x_iters is zero-initialised
        }
        __syncthreads();
        for (int ry = threadIdx.y; ry < 32; ry += 8) {
            const int r = r0 + ry, c = c0 + threadIdx.x;
            if (r < n && c < ws) o[(long long)r * ws + c] = tile[threadIdx.x][ry];
        }
        __syncthreads();
    }
}

// one CTA per active instance: vec[i] = 1 if p > hi, 0 if p < lo, else -1; num = #fixed, or 0 when #fixed <= min_fix
__global__ void lp_threshold_kernel(BatchView bv, const int *__restrict__ active, const long long *__restrict__ row_off,
                                    const float *__restrict__ scores, double hi, double lo, int min_fix, double *__restrict__ vec,
                                    long long *__restrict__ off_vec, int *__restrict__ num, const L2fMeta *__restrict__ meta) {
    __shared__ int s_cnt;
    if (meta && (long long)blockIdx.x >= meta->n_active) return;
    const int inst = active[blockIdx.x];
    const int n = bv.st[inst].n;
    const float *p = scores + row_off[blockIdx.x];
    double *v = vec + bv.off_n[inst];
    if (threadIdx.x == 0) { s_cnt = 0; off_vec[inst] = bv.off_n[inst]; }
    __syncthreads();
    int c = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double pi = (double)p[i];
        double f = -1.0;
        if (pi > hi) { f = 1.0; c++; } else if (pi < lo) { f = 0.0; c++; }
        v[i] = f;
    }
    atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) num[inst] = (s_cnt <= min_fix) ? 0 : s_cnt;
}

// Optional feasibility guard for fix-to-one decisions (NOT in the reference, whose deter_fix_2 trusts the policy; off by
// default).  One CTA per active instance, after lp_threshold_kernel: a variable proposed for x = 1 is kept only if, in EVERY
// constraint row it touches, (i) the row still has capacity for it (f_i - E_ij >= 0 with the current right-hand side, which
// already accounts for earlier fixes, LP.cpp:1276-1278) and (ii) it is the highest-scored proposal of that row (ties: lower
// index).  Two accepted variables can then never share a row, so the fixed part of the solution satisfies E x <= f by
// construction; losers fall back to "keep" (-1) and may be proposed again in a later window.  Proposals for x = 0 are always
// feasible for <= constraints with non-negative E and are left alone.  The <= min_fix rule is applied again afterwards.
__global__ void __launch_bounds__(256) lp_guard_kernel(BatchView bv, const int *__restrict__ active, const long long *__restrict__ row_off,
                                                       const float *__restrict__ scores, int min_fix, double *__restrict__ vec,
                                                       int *__restrict__ num, const L2fMeta *__restrict__ meta) {
    extern __shared__ unsigned long long s_best[];       // [m] best proposal key per row
    __shared__ int s_cnt;
    if (meta && (long long)blockIdx.x >= meta->n_active) return;
    const int inst = active[blockIdx.x];
    const InstState *st = bv.st + inst;
    const int n = st->n, m = st->m;
    if (num[inst] == 0) return;                          // nothing proposed (or already below the min_fix rule)
    const CsrLayout PL = csr_layout(st->n0, st->m0, st->nnz0);
    const unsigned char *pat = bv.csr + bv.off_csr[inst];
    const u16 *colptr = reinterpret_cast<const u16 *>(pat + PL.o_colptr);
    const u16 *rowidx = reinterpret_cast<const u16 *>(pat + PL.o_rowidx);
    const long long ov = bv.off_val ? bv.off_val[inst] : 0;
    const bool unit = st->unit != 0;
    const float *p = scores + row_off[blockIdx.x];
    double *v = vec + bv.off_n[inst];
    const double *f = bv.f + bv.off_m[inst];
    for (int i = threadIdx.x; i < m; i += blockDim.x) s_best[i] = 0ull;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    auto key = [&](int j) { return ((unsigned long long)__float_as_uint(p[j]) << 32) | (unsigned long long)(0xffffffffu - (unsigned)j); };   // scores are in (0, 1): bit order = value order
    for (int j = threadIdx.x; j < n; j += blockDim.x)
        if (v[j] == 1.0) { const unsigned long long k = key(j); for (int q = colptr[j]; q < colptr[j + 1]; ++q) atomicMax(&s_best[rowidx[q]], k); }
    __syncthreads();
    int c = 0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        double t = v[j];
        if (t == 1.0) {
            const unsigned long long k = key(j);
            bool ok = true;
            for (int q = colptr[j]; q < colptr[j + 1] && ok; ++q) {
                const int i = rowidx[q];
                const double e = unit ? 1.0 : bv.val_c[ov + q];
                ok = (s_best[i] == k) && (f[i] - e >= 0.0);
            }
            if (!ok) { t = -1.0; v[j] = t; }
        }
        if (t == 1.0 || t == 0.0) c++;
    }
    atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) num[inst] = (s_cnt <= min_fix) ? 0 : s_cnt;
}

}  // namespace lpb
