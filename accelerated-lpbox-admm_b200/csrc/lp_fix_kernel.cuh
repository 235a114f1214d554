// Device-side early fixing: the fix / compact prologue of ADMM_lp_iters_l2f (LP.cpp:1124-1335) + update_expression
// (LP.cpp:2289-2404), one CTA per instance, no host round trip.
//
// Given the fix vector over the CURRENT variables (1 / 0 = fix to that value, anything else = keep), the kernel
//   * splits the columns of E into kept (order preserved, renumbered) and fixed ones         LP.cpp:1135-1183
//   * appends (original id, value) of the fixed variables to ret_idx / ret_val, compacts left_idx   :1192-1206
//   * gathers x, y1, y2, z1, z2, b over the kept variables (y3, z4 keep length m)            :1222-1231
//   * fix_obj = b2 . x2 (Eigen reduction order), sum_fix_obj += fix_obj                      :1237-1249
//   * f <- f - E2 x2 (row-sequential, ascending column)                                      :1276-1278
//   * rebuilds both orientations of the pattern in place and the operator at the CURRENT rho :1295-1329
// All arithmetic follows the same ordering rules as lp_kernels.cuh (parity mode).
#pragma once
#include "lp_kernels.cuh"

namespace lpb {

constexpr int FIX_T = 256;

// in-place exclusive scan of data[0..L) (shared memory), all FIX_T threads participate; returns the total
__device__ __forceinline__ int block_exscan(int *data, int L, int *s_part) {
    const int tid = threadIdx.x;
    const int C = (L + FIX_T - 1) / FIX_T;
    const int beg = min(tid * C, L), end = min(beg + C, L);
    int sum = 0;
    for (int i = beg; i < end; ++i) sum += data[i];
    s_part[tid] = sum;
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int t = 0; t < FIX_T; ++t) { int v = s_part[t]; s_part[t] = run; run += v; }
        s_part[FIX_T] = run;
    }
    __syncthreads();
    int run = s_part[tid];
    for (int i = beg; i < end; ++i) { int v = data[i]; data[i] = run; run += v; }
    __syncthreads();
    return s_part[FIX_T];
}

// shared memory: int kidx[np+2] | int cnt[max(np,mp)+2] | int part[FIX_T+2] | int pold[np+2] | int s_w[80] |
//                double stage[max(np, val_elems)+2] | copy of the (old) CSR/CSC blob
__host__ __device__ inline size_t fix_ints(int np, int mp) {
    int L = (np > mp ? np : mp) + 2;
    size_t ints = (size_t)(np + 2) + (size_t)L + FIX_T + 2 + (size_t)(np + 2) + 80;
    return (ints + 3) & ~(size_t)3;
}
__host__ __device__ inline size_t fix_smem_bytes(int np, int mp, int csr_bytes, int val_elems) {
    size_t dbl = (size_t)(np > val_elems ? np : val_elems) + 2;
    return fix_ints(np, mp) * 4 + dbl * 8 + (size_t)csr_bytes + 16;
}

__global__ void __launch_bounds__(FIX_T)
lp_fix_kernel(BatchView bv, Params pr, const double *__restrict__ vec, const long long *__restrict__ off_vec,
              const int *__restrict__ num, int skip_done, int np, int mp, int csr_bytes, int val_elems) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int inst = blockIdx.x;
    InstState *st = bv.st + inst;
    if (skip_done && st->done) return;
    const int tid = threadIdx.x;
    const int n = st->n, m = st->m;
    const int fn = num ? num[inst] : 0;
    if (fn == 0 || n == 0) {                       // x_iters = Zero(n - 0, 500)  (LP.cpp:1113)
        if (tid == 0) { st->xit_rows = n; st->xit_cols = 0; st->norm_small = 0; }
        return;
    }
    const int L = (np > mp ? np : mp) + 2;
    const size_t nints = fix_ints(np, mp);
    int *kidx = reinterpret_cast<int *>(smem_raw);
    int *cnt = kidx + (np + 2);
    int *part = cnt + L;
    int *pold = part + FIX_T + 2;
    int *s_w = pold + (np + 2);
    double *stage = reinterpret_cast<double *>(smem_raw + nints * 4);
    unsigned char *spat = reinterpret_cast<unsigned char *>(stage + (size_t)(np > val_elems ? np : val_elems) + 2);

    const long long on = bv.off_n[inst], om = bv.off_m[inst], ov = bv.off_val ? bv.off_val[inst] : 0;
    const bool unit = st->unit != 0;
    const double *v = vec + off_vec[inst];
    const CsrLayout PL = csr_layout(st->n0, st->m0, st->nnz0);
    unsigned char *gpat = bv.csr + bv.off_csr[inst];
    const EllLayout EL = ell_layout(st->n0, st->m0, st->rcap, st->ccap);
    unsigned char *gell = bv.pat + bv.off_pat[inst];

    // stage the (old) pattern
    for (int k = tid; k < PL.bytes / 16; k += FIX_T)
        reinterpret_cast<uint4 *>(spat)[k] = reinterpret_cast<const uint4 *>(gpat)[k];
    // 1. classify (LP.cpp:1135-1150): 1 -> fix to 1, 0 -> fix to 0, else keep
    for (int i = tid; i < n; i += FIX_T) { double t = v[i]; kidx[i] = (t == 1.0 || t == 0.0) ? 0 : 1; }
    __syncthreads();
    const int k_tot = block_exscan(kidx, n, part);     // kidx[i] = new id of a kept variable
    const int j_tot = n - k_tot;
    const u16 *rowptr = reinterpret_cast<const u16 *>(spat + PL.o_rowptr);
    const u16 *colptr = reinterpret_cast<const u16 *>(spat + PL.o_colptr);
    const u16 *colidx = reinterpret_cast<const u16 *>(spat + PL.o_colidx);
    const u16 *rowidx = reinterpret_cast<const u16 *>(spat + PL.o_rowidx);
    auto is_fixed = [&](int i) { double t = v[i]; return t == 1.0 || t == 0.0; };

    // 2. ret_idx / ret_val append, left_idx compaction (:1192-1206)
    const int n_ret = st->n_ret;
    for (int i = tid; i < n; i += FIX_T) cnt[i] = bv.left_idx[on + i];
    __syncthreads();
    for (int i = tid; i < n; i += FIX_T) {
        if (is_fixed(i)) { int q = n_ret + (i - kidx[i]); bv.ret_idx[on + q] = cnt[i]; bv.ret_val[on + q] = v[i]; }
        else bv.left_idx[on + kidx[i]] = cnt[i];
    }
    __syncthreads();
    if (k_tot == 0) {                                  // :1212-1217
        if (tid == 0) {
            st->n = 0; st->nnz = 0; st->n_ret = n_ret + j_tot; st->fix_sum += j_tot; st->status = STOP_EMPTY; st->last_ret = 1; st->done = 1;
            st->xit_rows = 0; st->xit_cols = 0;
        }
        return;
    }
    // 4. fix_obj = b2.dot(x2) in Eigen order (:1237): products in fixed-variable order
    for (int i = tid; i < n; i += FIX_T)
        if (is_fixed(i)) stage[i - kidx[i]] = dM(bv.b[on + i], v[i]);
    __syncthreads();
    if (tid < 32) {
        double fo = warp_redux_eigen<1>(stage, 0, j_tot);
        if (tid == 0) {
            st->fix_obj = fo; st->prev_sum = st->sum_fix_obj; st->sum_fix_obj = dA(st->sum_fix_obj, fo);   // :1247-1248
            st->prev_obj = st->cur_obj;
        }
    }
    // 5. f1 = f - E2 x2 (:1276-1278) on the old row-compressed pattern
    for (int i = tid; i < m; i += FIX_T) {
        double acc = 0.0;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            int c = colidx[k];
            if (is_fixed(c)) acc = dA(acc, unit ? v[c] : dM(bv.val_r[ov + k], v[c]));
        }
        bv.f[om + i] = dS(bv.f[om + i], acc);
    }
    __syncthreads();
    // 6. gathers (:1222-1231)
    double *vecs[6] = {bv.x, bv.y1, bv.y2, bv.z1, bv.z2, bv.b};
    for (int a = 0; a < 6; ++a) {
        double *g = vecs[a] + on;
        for (int i = tid; i < n; i += FIX_T) stage[i] = g[i];
        __syncthreads();
        for (int i = tid; i < n; i += FIX_T) if (!is_fixed(i)) g[kidx[i]] = stage[i];
        __syncthreads();
    }
    // 7. x_sol.norm() < 1e-3 -> ret = 1 (:1223)
    for (int i = tid; i < k_tot; i += FIX_T) { double t = bv.x[on + i]; stage[i] = dM(t, t); }
    __syncthreads();
    if (tid < 32) {
        double s2 = warp_redux_eigen<1>(stage, 0, k_tot);
        if (tid == 0) st->norm_small = (sqrt(s2) < 1e-3) ? 1 : 0;
    }
    __syncthreads();
    // 8. pattern compaction (:1135-1183).  Row-compressed orientation: keep entries of kept columns, renumbered.
    u16 *g_rowptr = reinterpret_cast<u16 *>(gpat + PL.o_rowptr);
    u16 *g_colptr = reinterpret_cast<u16 *>(gpat + PL.o_colptr);
    u16 *g_colidx = reinterpret_cast<u16 *>(gpat + PL.o_colidx);
    u16 *g_rowidx = reinterpret_cast<u16 *>(gpat + PL.o_rowidx);
    for (int i = tid; i < m; i += FIX_T) {
        int c = 0;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) c += is_fixed(colidx[k]) ? 0 : 1;
        cnt[i] = c;
    }
    __syncthreads();
    const int nnz_new = block_exscan(cnt, m, part);
    if (!unit) {
        for (int k = tid; k < st->nnz; k += FIX_T) stage[k] = bv.val_r[ov + k];
        __syncthreads();
    }
    for (int i = tid; i < m; i += FIX_T) {
        int q = cnt[i];
        g_rowptr[i] = (u16)q;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            int c = colidx[k];
            if (!is_fixed(c)) { g_colidx[q] = (u16)kidx[c]; if (!unit) bv.val_r[ov + q] = stage[k]; q++; }
        }
    }
    if (tid == 0) g_rowptr[m] = (u16)nnz_new;
    __syncthreads();
    // column-compressed orientation: drop fixed columns
    for (int i = tid; i < n; i += FIX_T) cnt[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += FIX_T) if (!is_fixed(i)) cnt[kidx[i]] = colptr[i + 1] - colptr[i];
    __syncthreads();
    block_exscan(cnt, k_tot, part);
    if (!unit) {
        for (int k = tid; k < st->nnz; k += FIX_T) stage[k] = bv.val_c[ov + k];
        __syncthreads();
    }
    for (int i = tid; i < n; i += FIX_T) {
        if (!is_fixed(i)) {
            int q = cnt[kidx[i]];
            g_colptr[kidx[i]] = (u16)q;
            for (int k = colptr[i]; k < colptr[i + 1]; ++k) { g_rowidx[q] = rowidx[k]; if (!unit) bv.val_c[ov + q] = stage[k]; q++; }
        }
    }
    if (tid == 0) g_colptr[k_tot] = (u16)nnz_new;
    __syncthreads();
    // column work assignment: stable filter of the old (length-sorted) slot order, renumbered
    {
        u16 *g_cperm = reinterpret_cast<u16 *>(gell + EL.o_cperm);
        for (int s2 = tid; s2 < n; s2 += FIX_T) { int c = g_cperm[s2]; pold[s2] = c; cnt[s2] = is_fixed(c) ? 0 : 1; }
        __syncthreads();
        block_exscan(cnt, n, part);
        for (int s2 = tid; s2 < n; s2 += FIX_T) { int c = pold[s2]; if (!is_fixed(c)) g_cperm[cnt[s2]] = (u16)kidx[c]; }
    }
    __syncthreads();
    // padded sliced-ELL image of the compacted pattern (rows keep their slots; lengths only shrink); cnt / pold are free now
    build_ell_image(bv, inst, st, k_tot, m, g_rowptr, g_colidx, g_colptr, g_rowidx, unit, ov, reinterpret_cast<u16 *>(cnt),
                    reinterpret_cast<u16 *>(pold), s_w);
    // 9. update_expression with the current rho (:1329, :2289-2404) on the new column-compressed pattern
    const double rho1 = st->rho1, rho2 = st->rho2, rho4 = st->rho4;
    const double D = dA(0.0, dA(rho1, rho2));
    for (int j = tid; j < k_tot; j += FIX_T) {
        double e = 0.0;
        const int kb = g_colptr[j], ke = g_colptr[j + 1];
        for (int k = kb; k < ke; ++k) {
            if (unit) e = dA(e, 1.0);
            else { double t = bv.val_c[ov + k]; if (t != 0.0) e = dA(e, dM(t, t)); }
        }
        bv.Esq[on + j] = e;
        bv.Pd[on + j] = dA(D, dM(rho4, e));
    }
    if (!unit) {
        const long long ec = bv.off_evc[inst];
        for (int k = tid; k < 32 * st->ccap; k += FIX_T) bv.r4v[ec + k] = dM(rho4, bv.ev_c[ec + k]);
    }
    if (tid == 0) {
        st->D = D; st->r4s = dM(rho4, 1.0);
        st->n = k_tot; st->nnz = nnz_new; st->n_ret = n_ret + j_tot; st->fix_sum += j_tot;
        st->xit_rows = k_tot; st->xit_cols = 0;
    }
}

}  // namespace lpb
