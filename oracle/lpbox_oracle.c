/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's Lp-Box ADMM hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library.  It is the checker the CUDA path is compared with, never a fallback for it.
 *
 * Restates, in plain C (no Eigen), the arithmetic of
 *   LP.cpp  = LinerProgramming/LinearProgramming/cython_solver/LPboxADMMsolver.cpp
 *   SEG.cpp = Segmentation/Segmentation/cython/src/LPboxADMMsolver.cpp
 * following the operation-order rules of SURVEY.md §8c / Appendix A:
 *   - fp64, no FMA (build with -ffp-contract=off, no -march),
 *   - SpMV: one accumulator per output row, stored entries in ascending inner index, starting from 0,
 *   - reductions in Eigen's SSE2 order: two 2-lane packets, i.e. four interleaved sequential chains,
 *   - expression association exactly as written in the C++ source.
 * Parity status: PINNED against the reference's own compiled Eigen build (oracle/_ref/liblpbox_solver.so,
 * driven by oracle/ref_harness.py) -- see tests/test_oracle_vs_reference.py -- and against the golden
 * vectors under tests/golden/ generated from that binary.
 *
 * Every function cites the reference lines it follows.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define LPO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------
 * Eigen BLAS-1 semantics
 * ---------------------------------------------------------------------------------------------- */

/* Eigen redux_impl<Func, Evaluator, LinearVectorizedTraversal, NoUnrolling> with Packet2d, alignedStart=0
 * (call sites: LP.cpp:277,288,300,306,311,323,425,455,931-933).  `get(i)` yields element i of the
 * expression being reduced. */
#define LPO_REDUX_BODY(GET)                                                       \
    long a2 = (n / 4) * 4, a1 = (n / 2) * 2;                                       \
    double res;                                                                    \
    if (a1) {                                                                      \
        double p00 = GET(0), p01 = GET(1);                                         \
        if (a1 > 2) {                                                              \
            double p10 = GET(2), p11 = GET(3);                                     \
            for (long i = 4; i < a2; i += 4) {                                     \
                p00 = p00 + GET(i);     p01 = p01 + GET(i + 1);                    \
                p10 = p10 + GET(i + 2); p11 = p11 + GET(i + 3);                    \
            }                                                                      \
            p00 = p00 + p10; p01 = p01 + p11;                                      \
            if (a1 > a2) { p00 = p00 + GET(a2); p01 = p01 + GET(a2 + 1); }         \
        }                                                                          \
        res = p00 + p01;                                                           \
        for (long i = a1; i < n; ++i) res = res + GET(i);                          \
    } else {                                                                       \
        if (n == 0) return 0.0;                                                    \
        res = GET(0);                                                              \
        for (long i = 1; i < n; ++i) res = res + GET(i);                           \
    }                                                                              \
    return res;

LPO_API double lpo_sum(const double *v, long n) {
#define G(i) (v[i])
    LPO_REDUX_BODY(G)
#undef G
}
LPO_API double lpo_dot(const double *a, const double *b, long n) {
#define G(i) (a[i] * b[i])
    LPO_REDUX_BODY(G)
#undef G
}
LPO_API double lpo_sqnorm(const double *a, long n) { return lpo_dot(a, a, n); }
LPO_API double lpo_norm(const double *a, long n) { return sqrt(lpo_sqnorm(a, n)); }
/* ||a-b||: the difference is an expression evaluated per coefficient inside the redux (LP.cpp:932-933) */
static double sqnorm_diff(const double *a, const double *b, long n) {
#define G(i) ((a[i] - b[i]) * (a[i] - b[i]))
    LPO_REDUX_BODY(G)
#undef G
}
/* ||a-c|| with scalar c (project_shifted_Lp_ball first materialises y = x-0.5, then y.norm(); LP.cpp:424-425) */

/* ------------------------------------------------------------------------------------------------
 * Compressed sparse storage.  `ptr/idx/val` are "outer/inner" arrays; we keep both orientations of E.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int nouter, ninner;
    int *ptr;
    int *idx;
    double *val;
} csx_t;

static void csx_free(csx_t *a) {
    free(a->ptr); free(a->idx); free(a->val);
    memset(a, 0, sizeof(*a));
}
static void csx_alloc(csx_t *a, int nouter, int ninner, int nnz) {
    a->nouter = nouter; a->ninner = ninner;
    a->ptr = (int *)calloc((size_t)nouter + 1, sizeof(int));
    a->idx = (int *)malloc(sizeof(int) * (size_t)(nnz > 0 ? nnz : 1));
    a->val = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
}
static void csx_copy(csx_t *d, const csx_t *s) {
    int nnz = s->ptr[s->nouter];
    csx_alloc(d, s->nouter, s->ninner, nnz);
    memcpy(d->ptr, s->ptr, sizeof(int) * ((size_t)s->nouter + 1));
    memcpy(d->idx, s->idx, sizeof(int) * (size_t)nnz);
    memcpy(d->val, s->val, sizeof(double) * (size_t)nnz);
}
/* transpose keeping ascending inner order (Eigen transpose of a compressed matrix does the same) */
static void csx_transpose(csx_t *d, const csx_t *s) {
    int nnz = s->ptr[s->nouter];
    csx_alloc(d, s->ninner, s->nouter, nnz);
    for (int k = 0; k < nnz; ++k) d->ptr[s->idx[k] + 1]++;
    for (int i = 0; i < d->nouter; ++i) d->ptr[i + 1] += d->ptr[i];
    int *pos = (int *)malloc(sizeof(int) * ((size_t)d->nouter + 1));
    memcpy(pos, d->ptr, sizeof(int) * ((size_t)d->nouter + 1));
    for (int o = 0; o < s->nouter; ++o)
        for (int k = s->ptr[o]; k < s->ptr[o + 1]; ++k) {
            int q = pos[s->idx[k]]++;
            d->idx[q] = o; d->val[q] = s->val[k];
        }
    free(pos);
}
/* y = M v, M given by its row-compressed form: y_i = ((0 + v_i1 x_j1) + v_i2 x_j2) + ...  (LP.cpp:102-108;
 * Eigen's ColMajor scatter kernel gives the same per-row order, SURVEY.md §8c rule 2) */
static void spmv_rows(const csx_t *rows, const double *val, const double *x, double *y) {
    for (int i = 0; i < rows->nouter; ++i) {
        double acc = 0.0;
        for (int k = rows->ptr[i]; k < rows->ptr[i + 1]; ++k) acc = acc + val[k] * x[rows->idx[k]];
        y[i] = acc;
    }
}
LPO_API void lpo_spmv_csr(int nrows, const int *rowptr, const int *colidx, const double *val, const double *x, double *y) {
    csx_t r = {nrows, 0, (int *)rowptr, (int *)colidx, (double *)val};
    spmv_rows(&r, val, x, y);
}

/* ------------------------------------------------------------------------------------------------
 * Solver object
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    /* hyper-parameters, LP.h:115-146 */
    double stop_threshold, std_threshold;
    int max_iters;
    double initial_rho;
    int rho_change_step;
    double gamma_val, learning_fact, history_size, projection_lp, gamma_factor, pcg_tol;
    int pcg_maxiters;
    /* behaviour switches: which loop variant of the reference is being restated */
    int guard_first_iter;  /* LP.cpp:934 `iter != iter_start` (plain LP loop only) */
    int alpha_bailout;     /* LP.cpp:307  `if(alpha < 0) return -1` (LP variant only) */
    int quadratic;         /* unconstrained form with explicit A (SEG.cpp) */
    int has_ineq;

    /* problem (current = possibly compacted) */
    int n, m, org_n;
    csx_t Ec;   /* E column-compressed (outer = column), as LP.h:17 */
    csx_t Er;   /* E row-compressed (outer = row) */
    csx_t Eorg; /* org_E_ptr, column-compressed */
    double *b, *f;
    csx_t A;    /* quadratic term (row-compressed, symmetric), SEG only */
    double cconst;

    /* state, LP.h:199-262 */
    double *x, *y1, *y2, *z1, *z2, *y3, *z4, *best_sol;
    double *D;      /* diagonal of _2A_plus_rho1_rho2 (LP) */
    double *Pd;     /* diagonal of preconditioner_diag_mat */
    double *invd;   /* DiagonalPreconditioner::m_invdiag */
    double *Esq;
    double *r4val;  /* values of rho4_E_transpose, stored in Ec (column-of-E) order */
    csx_t M;        /* SEG: temp_mat = 2A + (rho1+rho2) I, explicit (row-compressed) */
    int *Mdiag;     /* position of the diagonal entry of each row of M */
    double rho1, rho2, rho3, rho4, prho1, prho2, prho3, prho4, gamma, ratio;
    int rhoUpdated;
    double *obj_list; long obj_len, obj_cap;
    double std_obj, cur_obj, best_bin_obj, sum_fix_obj, fix_obj, prev_obj, prev_sum;
    int iter;
    /* early fixing */
    int *left_idx; int n_left_idx;
    int *ret_idx; double *ret_val; int n_ret;   /* ret_idx_prev / ret_val_prev */
    int fix_sum;
    double *x_iters; int xit_rows, xit_cols;    /* DenseMatrix x_iters (rows x 500), column-major like Eigen */
    /* counters (not in the reference; for metrics) */
    long cg_iters_total, admm_iters_total;
    int last_cg_iters;
    /* scratch */
    double *t_n0, *t_n1, *t_n2, *t_n3, *t_m0, *t_m1;
    int cap_n, cap_m;
} lpo_solver;

LPO_API lpo_solver *lpo_create(void) {
    lpo_solver *s = (lpo_solver *)calloc(1, sizeof(lpo_solver));
    s->std_obj = 1.0;   /* LP.h:214 */
    s->rhoUpdated = 1;  /* LP.h:208 */
    return s;
}

static void free_vecs(lpo_solver *s) {
    double **v[] = {&s->x, &s->y1, &s->y2, &s->z1, &s->z2, &s->y3, &s->z4, &s->best_sol, &s->D, &s->Pd, &s->invd,
                    &s->Esq, &s->r4val, &s->t_n0, &s->t_n1, &s->t_n2, &s->t_n3, &s->t_m0, &s->t_m1};
    for (size_t i = 0; i < sizeof(v) / sizeof(v[0]); ++i) { free(*v[i]); *v[i] = NULL; }
}

LPO_API void lpo_destroy(lpo_solver *s) {
    if (!s) return;
    free_vecs(s);
    csx_free(&s->Ec); csx_free(&s->Er); csx_free(&s->Eorg); csx_free(&s->A); csx_free(&s->M);
    free(s->Mdiag); free(s->b); free(s->f); free(s->obj_list); free(s->left_idx); free(s->ret_idx); free(s->ret_val);
    free(s->x_iters);
    free(s);
}

/* LP.cpp:491-507 */
LPO_API void lpo_params_lp(lpo_solver *s) {
    s->stop_threshold = 1e-4; s->std_threshold = 1e-12; s->gamma_val = 1.6; s->gamma_factor = 0.95;
    s->rho_change_step = 25; s->max_iters = (int)2e4; s->initial_rho = 25; s->history_size = 10;
    s->learning_fact = 1 + 1.0 / 100; s->pcg_tol = 1e-3; s->pcg_maxiters = (int)1e3; s->projection_lp = 2;
}
/* SEG.cpp:659-672 */
LPO_API void lpo_params_seg(lpo_solver *s) {
    s->std_threshold = 1e-6; s->gamma_val = 1.0; s->gamma_factor = 0.99; s->initial_rho = 5;
    s->learning_fact = 1 + 3.0 / 100; s->history_size = 5; s->rho_change_step = 5; s->stop_threshold = 1e-3;
    s->max_iters = (int)1e4; s->projection_lp = 2; s->pcg_tol = 1e-3; s->pcg_maxiters = (int)1e3;
}
/* header setters LP.h:511-575 (never called from Cython; used by the parity sweeps) */
LPO_API void lpo_set_params(lpo_solver *s, double stop_threshold, double std_threshold, int max_iters, double initial_rho,
                            int rho_change_step, double gamma_val, double learning_fact, double history_size,
                            double projection_lp, double gamma_factor, double pcg_tol, int pcg_maxiters) {
    s->stop_threshold = stop_threshold; s->std_threshold = std_threshold; s->max_iters = max_iters;
    s->initial_rho = initial_rho; s->rho_change_step = rho_change_step; s->gamma_val = gamma_val;
    s->learning_fact = learning_fact; s->history_size = history_size; s->projection_lp = projection_lp;
    s->gamma_factor = gamma_factor; s->pcg_tol = pcg_tol; s->pcg_maxiters = pcg_maxiters;
}
LPO_API void lpo_set_variant(lpo_solver *s, int guard_first_iter, int alpha_bailout) {
    s->guard_first_iter = guard_first_iter; s->alpha_bailout = alpha_bailout;
}

/* problem in: E column-compressed (what readFile builds, LP.cpp:2416-2444,:2508-2533), b as the solver sees it
 * (readFile negates the bid prices, LP.cpp:2520) and f (ones, LP.cpp:2522). */
LPO_API int lpo_set_problem_csc(lpo_solver *s, int m, int n, const int *colptr, const int *rowidx, const double *val,
                                const double *b, const double *f) {
    csx_free(&s->Ec); csx_free(&s->Er); csx_free(&s->Eorg);
    free(s->b); free(s->f);
    int nnz = colptr[n];
    csx_alloc(&s->Ec, n, m, nnz);
    memcpy(s->Ec.ptr, colptr, sizeof(int) * ((size_t)n + 1));
    memcpy(s->Ec.idx, rowidx, sizeof(int) * (size_t)nnz);
    memcpy(s->Ec.val, val, sizeof(double) * (size_t)nnz);
    for (int j = 0; j < n; ++j)
        for (int k = colptr[j] + 1; k < colptr[j + 1]; ++k)
            if (rowidx[k] <= rowidx[k - 1]) return -1; /* must be strictly ascending (setFromTriplets output) */
    csx_transpose(&s->Er, &s->Ec);
    csx_copy(&s->Eorg, &s->Ec);
    s->b = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    s->f = (double *)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1));
    memcpy(s->b, b, sizeof(double) * (size_t)n);
    memcpy(s->f, f, sizeof(double) * (size_t)m);
    s->n = n; s->m = m; s->org_n = n; s->has_ineq = 1; s->quadratic = 0;
    return 0;
}

static void alloc_state(lpo_solver *s, int n, int m) {
    free_vecs(s);
    size_t nn = (size_t)(n > 0 ? n : 1), mm = (size_t)(m > 0 ? m : 1);
    double **nv[] = {&s->x, &s->y1, &s->y2, &s->z1, &s->z2, &s->best_sol, &s->D, &s->Pd, &s->invd, &s->Esq,
                     &s->t_n0, &s->t_n1, &s->t_n2, &s->t_n3};
    for (size_t i = 0; i < sizeof(nv) / sizeof(nv[0]); ++i) *nv[i] = (double *)calloc(nn, sizeof(double));
    double **mv[] = {&s->y3, &s->z4, &s->t_m0, &s->t_m1};
    for (size_t i = 0; i < sizeof(mv) / sizeof(mv[0]); ++i) *mv[i] = (double *)calloc(mm, sizeof(double));
    s->cap_n = n; s->cap_m = m;
}

/* update_expression, LP.cpp:2289-2404 */
static void update_expression(lpo_solver *s) {
    int n = s->n;
    int nnz = s->Ec.ptr[n];
    free(s->r4val);
    s->r4val = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    /* :2292-2293  rho4_E_transpose = rho4 * E_transpose (scaled values stored) */
    for (int k = 0; k < nnz; ++k) s->r4val[k] = s->rho4 * s->Ec.val[k];
    /* :2339-2343  zero diagonal, then += rho1 + rho2 */
    for (int j = 0; j < n; ++j) s->D[j] = 0.0 + (s->rho1 + s->rho2);
    /* :2351 preconditioner_diag_mat = _2A_plus_rho1_rho2 */
    for (int j = 0; j < n; ++j) s->Pd[j] = s->D[j];
    /* :2379-2390 Esq_j = sequential sum of squares over column j, skipping explicit zeros */
    for (int j = 0; j < n; ++j) {
        double e = 0.0;
        for (int k = s->Ec.ptr[j]; k < s->Ec.ptr[j + 1]; ++k)
            if (s->Ec.val[k] != 0.0) e += s->Ec.val[k] * s->Ec.val[k];
        s->Esq[j] = e;
    }
    /* :2391 */
    for (int j = 0; j < n; ++j) s->Pd[j] = s->Pd[j] + s->rho4 * s->Esq[j];
}

/* ADMM_lp_iters_init, LP.cpp:489-763 */
LPO_API int lpo_lp_init(lpo_solver *s) {
    lpo_params_lp(s);
    int n = s->n, m = s->m;
    alloc_state(s, n, m);
    free(s->left_idx);
    s->left_idx = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) { s->left_idx[i] = i; s->x[i] = 1.0; }      /* :582-586 */
    s->n_left_idx = n;
    s->org_n = n; s->fix_sum = 0;
    free(s->ret_idx); free(s->ret_val); s->ret_idx = NULL; s->ret_val = NULL; s->n_ret = 0; /* ret_*_prev are empty */
    s->fix_obj = 0; s->sum_fix_obj = 0;
    s->rho1 = s->rho2 = s->rho3 = s->rho4 = s->initial_rho;                 /* :629-636 */
    s->prho1 = s->prho2 = s->prho3 = s->prho4 = s->initial_rho;
    s->gamma = s->gamma_val;
    memcpy(s->y1, s->x, sizeof(double) * (size_t)n);                         /* :714-715 */
    memcpy(s->y2, s->x, sizeof(double) * (size_t)n);
    spmv_rows(&s->Er, s->Er.val, s->x, s->t_m0);                             /* :720 y3 = f - E x */
    for (int i = 0; i < m; ++i) s->y3[i] = s->f[i] - s->t_m0[i];
    memcpy(s->best_sol, s->x, sizeof(double) * (size_t)n);
    s->best_bin_obj = lpo_dot(s->b, s->x, n);                                /* :726 compute_cost_lp(x_sol,b) = b.dot(x) */
    s->guard_first_iter = 1; s->alpha_bailout = 1;
    s->obj_len = 0; s->std_obj = 1.0; s->cur_obj = 0; s->rhoUpdated = 1; s->iter = 0;
    s->cg_iters_total = 0; s->admm_iters_total = 0;
    return 1;
}

/* generic start used to mirror SEG.cpp:1384-1547 (ADMM_bqp prologue) for the inequality form with A = 0:
 * caller supplies x0 and hyper-parameters; operator matrices are built as in update_expression. */
LPO_API int lpo_generic_ineq_init(lpo_solver *s, const double *x0) {
    int n = s->n, m = s->m;
    alloc_state(s, n, m);
    memcpy(s->x, x0, sizeof(double) * (size_t)n);
    free(s->left_idx);
    s->left_idx = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) s->left_idx[i] = i;
    s->n_left_idx = n; s->org_n = n; s->n_ret = 0;
    s->rho1 = s->rho2 = s->rho3 = s->rho4 = s->initial_rho;
    s->prho1 = s->prho2 = s->prho3 = s->prho4 = s->initial_rho;
    s->gamma = s->gamma_val;
    memcpy(s->y1, s->x, sizeof(double) * (size_t)n);
    memcpy(s->y2, s->x, sizeof(double) * (size_t)n);
    spmv_rows(&s->Er, s->Er.val, s->x, s->t_m0);
    for (int i = 0; i < m; ++i) s->y3[i] = s->f[i] - s->t_m0[i];
    memcpy(s->best_sol, s->x, sizeof(double) * (size_t)n);
    s->best_bin_obj = 0.0 + lpo_dot(s->b, s->x, n);
    s->guard_first_iter = 0; s->alpha_bailout = 0;
    s->obj_len = 0; s->std_obj = 1.0; s->cur_obj = 0; s->rhoUpdated = 1; s->iter = 0;
    s->sum_fix_obj = 0; s->fix_sum = 0;
    s->cg_iters_total = 0; s->admm_iters_total = 0;
    return 1;
}

/* calculate_mat_expr_multiplication for {D} + {E, rho4 E^T}, LP.cpp:115-162:
 * result = D v (a diagonal *sparse* product: 0 + d_i v_i); temp = R4ET (E v); result += temp */
static void apply_op(lpo_solver *s, const double *v, double *out) {
    int n = s->n;
    double *t1 = s->t_m1;
    spmv_rows(&s->Er, s->Er.val, v, t1);
    for (int j = 0; j < n; ++j) {
        double acc = 0.0;
        for (int k = s->Ec.ptr[j]; k < s->Ec.ptr[j + 1]; ++k) acc = acc + s->r4val[k] * t1[s->Ec.idx[k]];
        double dv = 0.0 + s->D[j] * v[j];
        out[j] = dv + acc;
    }
}

/* _conjugate_gradient, LP.cpp:251-335.  Returns 1, or -1 on the alpha<0 bail-out (x as updated so far). */
static int pcg(lpo_solver *s, const double *rhs, double *x, int *iters_out) {
    int n = s->n;
    double *r = s->t_n1, *p = s->t_n2, *tmp = s->t_n3;
    double tol = s->pcg_tol;
    int maxIters = s->pcg_maxiters;
    apply_op(s, x, tmp);                                         /* :272 */
    for (int i = 0; i < n; ++i) r[i] = rhs[i] - tmp[i];           /* :273 */
    double rhsNorm2 = lpo_sqnorm(rhs, n);                         /* :277 */
    if (rhsNorm2 == 0) { memset(x, 0, sizeof(double) * (size_t)n); *iters_out = 0; return 1; } /* :279-284 */
    double threshold = tol * tol * rhsNorm2;                      /* :287 */
    if (!(threshold > DBL_MIN)) threshold = DBL_MIN;              /* numext::maxi(a,b) = a<b ? b : a */
    double r2 = lpo_sqnorm(r, n);                                 /* :288 */
    if (r2 < threshold) { *iters_out = 0; return 1; }             /* :290-295 */
    for (int i = 0; i < n; ++i) p[i] = s->invd[i] * r[i];         /* :297 precond.solve */
    double absNew = lpo_dot(r, p, n);                             /* :300 */
    int i = 0;
    while (i < maxIters) {
        apply_op(s, p, tmp);                                      /* :304 */
        double alpha = absNew / lpo_dot(p, tmp, n);               /* :306 */
        if (s->alpha_bailout && alpha < 0) { *iters_out = i; return -1; } /* :307 */
        for (int k = 0; k < n; ++k) x[k] = x[k] + alpha * p[k];   /* :308 */
        for (int k = 0; k < n; ++k) r[k] = r[k] - alpha * tmp[k]; /* :310 */
        r2 = lpo_sqnorm(r, n);                                    /* :311 */
        if (r2 < threshold) { i++; break; }                       /* :315-318 */
        /* :320-325  z = invdiag.*r ; absNew = r.z ; p = z + beta p   (z materialised in tmp) */
        for (int k = 0; k < n; ++k) tmp[k] = s->invd[k] * r[k];
        double absOld = absNew;
        absNew = lpo_dot(r, tmp, n);
        double beta = absNew / absOld;
        for (int k = 0; k < n; ++k) p[k] = tmp[k] + beta * p[k];
        i++;
    }
    *iters_out = i;
    return 1;
}

/* std_dev LP.cpp:358-377 + compute_std_obj :459-469 */
static double compute_std_obj(const double *list, long s, int history) {
    long begin = (s <= history) ? 0 : s - history, end = s;
    long size = end - begin;
    double mean = 0, sd = 0;
    for (long i = begin; i < end; ++i) mean += list[i];
    mean /= (double)size;
    for (long i = 0; i < size; ++i) sd += (list[begin + i] - mean) * (list[begin + i] - mean);
    sd /= (double)(size - 1);
    double r = (sd == 0) ? 0.0 : pow(sd, 1.0 / 2);
    return r / fabs(list[s - 1]);
}

static void push_obj(lpo_solver *s, double v) {
    if (s->obj_len == s->obj_cap) {
        s->obj_cap = s->obj_cap ? 2 * s->obj_cap : 1024;
        s->obj_list = (double *)realloc(s->obj_list, sizeof(double) * (size_t)s->obj_cap);
    }
    s->obj_list[s->obj_len++] = v;
}

/* One ADMM iteration of the inequality form (LP.cpp:801-1011 == :1346-1563 == SEG.cpp:1590-1760 restricted to
 * problem_type==inequality).  Returns 0 = continue, 1 = stopped by y1/y2 test, 2 = stopped by obj-std test,
 * 3 = CG alpha<0 bail-out (l2f loop only).  `record` != 0 stores x into x_iters column *cc (LP.cpp:1472-1475). */
static int ineq_iteration(lpo_solver *s, int iter, int iter_start, int l2f, int *cc) {
    int n = s->n, m = s->m;
    double *tn = s->t_n0;
    /* y1  :806-809 */
    for (int i = 0; i < n; ++i) {
        double t = s->x[i] + s->z1[i] / s->rho1;
        s->y1[i] = (t > 1) ? 1.0 : ((t < 0) ? 0.0 : t);
    }
    /* y2  :815-818, :423-428 */
    for (int i = 0; i < n; ++i) s->y2[i] = (s->x[i] + s->z2[i] / s->rho2) - 0.5;
    {
        double nrm = lpo_norm(s->y2, n);
        double c = pow((double)n, 1.0 / (int)s->projection_lp);
        for (int i = 0; i < n; ++i) s->y2[i] = s->y2[i] * c / (2 * nrm) + 0.5;
    }
    /* y3  :824-828 */
    spmv_rows(&s->Er, s->Er.val, s->x, s->t_m0);
    for (int i = 0; i < m; ++i) {
        double t = s->f[i] - s->t_m0[i] - s->z4[i] / s->rho4;
        s->y3[i] = (t < 0) ? 0.0 : t;
    }
    if (iter == 0) update_expression(s);                         /* :833 / :1380-1381 */
    /* :851-866 */
    if (iter != 0 && s->rhoUpdated) {
        double c12 = s->ratio * (s->prho1 + s->prho2);
        for (int i = 0; i < n; ++i) s->D[i] += c12;
        for (int i = 0; i < n; ++i) s->Pd[i] += c12;
        double c4 = s->ratio * s->prho4;
        for (int i = 0; i < n; ++i) s->Pd[i] += c4 * s->Esq[i];
        int nnz = s->Ec.ptr[n];
        for (int k = 0; k < nnz; ++k) s->r4val[k] = s->learning_fact * s->r4val[k];
    }
    /* rhs :872-878 */
    for (int i = 0; i < n; ++i) tn[i] = s->rho1 * s->y1[i] + s->rho2 * s->y2[i] - (s->b[i] + s->z1[i] + s->z2[i]);
    for (int i = 0; i < m; ++i) s->t_m0[i] = s->f[i] - s->y3[i];
    for (int j = 0; j < n; ++j) {
        double acc = 0.0;
        for (int k = s->Ec.ptr[j]; k < s->Ec.ptr[j + 1]; ++k) acc = acc + s->r4val[k] * s->t_m0[s->Ec.idx[k]];
        tn[j] += acc;
    }
    for (int j = 0; j < n; ++j) {
        double acc = 0.0;
        for (int k = s->Ec.ptr[j]; k < s->Ec.ptr[j + 1]; ++k) acc = acc + s->Ec.val[k] * s->z4[s->Ec.idx[k]];
        tn[j] -= acc;
    }
    /* :883-890  DiagonalPreconditioner::compute: invdiag = d != 0 ? 1/d : 1 */
    if (s->rhoUpdated) {
        for (int i = 0; i < n; ++i) s->invd[i] = (s->Pd[i] != 0.0) ? 1.0 / s->Pd[i] : 1.0;
        s->rhoUpdated = 0;
    }
    /* :892-896 / :1442-1458 */
    int cgit = 0;
    if (!l2f) {
        memcpy(s->x, s->y1, sizeof(double) * (size_t)n);
        pcg(s, tn, s->x, &cgit);       /* plain loop ignores the return value (:894) */
    } else {
        double *xt = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
        memcpy(xt, s->y1, sizeof(double) * (size_t)n);
        int cg = pcg(s, tn, xt, &cgit);
        if (cg == -1) { free(xt); s->cg_iters_total += cgit; return 3; }
        memcpy(s->x, xt, sizeof(double) * (size_t)n);
        free(xt);
    }
    s->last_cg_iters = cgit; s->cg_iters_total += cgit; s->admm_iters_total += 1;
    if (l2f && cc) {                                             /* :1472-1475 */
        if (*cc < s->xit_cols) for (int i = 0; i < n; ++i) s->x_iters[(size_t)(*cc) * (size_t)s->xit_rows + i] = s->x[i];
        (*cc)++;
    }
    /* duals :917-924 */
    {
        double g1 = s->gamma * s->rho1, g2 = s->gamma * s->rho2, g4 = s->gamma * s->rho4;
        for (int i = 0; i < n; ++i) s->z1[i] = s->z1[i] + g1 * (s->x[i] - s->y1[i]);
        for (int i = 0; i < n; ++i) s->z2[i] = s->z2[i] + g2 * (s->x[i] - s->y2[i]);
        spmv_rows(&s->Er, s->Er.val, s->x, s->t_m0);
        if (!l2f && s->guard_first_iter && iter == iter_start) /* :920-921: assignment, not accumulation */
            for (int i = 0; i < m; ++i) s->z4[i] = g4 * (s->t_m0[i] + s->y3[i] - s->f[i]);
        else
            for (int i = 0; i < m; ++i) s->z4[i] = s->z4[i] + g4 * (s->t_m0[i] + s->y3[i] - s->f[i]);
    }
    /* stop test 1 :931-949 */
    {
        double temp0 = lpo_norm(s->x, n);
        if (!(temp0 > 2.2204e-16)) temp0 = 2.2204e-16;  /* std::max(a,b) = a<b ? b : a */
        double c1 = sqrt(sqnorm_diff(s->x, s->y1, n)) / temp0;
        double c2 = sqrt(sqnorm_diff(s->x, s->y2, n)) / temp0;
        int guard = (!l2f && s->guard_first_iter) ? (iter != iter_start) : 1;
        if (c1 <= s->stop_threshold && c2 <= s->stop_threshold && guard) return 1;
    }
    /* rho schedule :951-970 */
    if ((iter + 1) % s->rho_change_step == 0) {
        s->prho1 = s->rho1; s->prho2 = s->rho2;
        s->rho1 = s->learning_fact * s->rho1; s->rho2 = s->learning_fact * s->rho2;
        s->prho4 = s->rho4; s->rho4 = s->learning_fact * s->rho4;
        double g = s->gamma * s->gamma_factor;
        s->gamma = (g < 1.0) ? 1.0 : g;
        s->rhoUpdated = 1;
        s->ratio = s->learning_fact - 1.0;
    }
    /* stop test 2 :972-995 */
    {
        double obj = lpo_dot(s->b, s->x, n);
        push_obj(s, obj);
        if ((double)s->obj_len >= s->history_size) s->std_obj = compute_std_obj(s->obj_list, s->obj_len, (int)s->history_size);
        if (s->std_obj <= s->std_threshold) return 2;
    }
    /* binary objective :1001-1011 */
    {
        for (int i = 0; i < n; ++i) tn[i] = (s->x[i] >= 0.5) ? 1.0 : 0.0;
        s->cur_obj = lpo_dot(s->b, tn, n);
        if (s->best_bin_obj >= s->cur_obj) { s->best_bin_obj = s->cur_obj; memcpy(s->best_sol, s->x, sizeof(double) * (size_t)n); }
    }
    return 0;
}

/* ADMM_lp_iters, LP.cpp:766-1095.  Return value as the reference: 1 only for the obj-std stop. */
LPO_API int lpo_lp_iters(lpo_solver *s, int iter_start, int iter_end) {
    int ret = 0, iter;
    for (iter = iter_start; iter < iter_end; ++iter) {
        int st = ineq_iteration(s, iter, iter_start, 0, NULL);
        if (st == 1) break;
        if (st == 2) { ret = 1; break; }
    }
    s->iter = iter; /* the plain loop uses a local `iter` (LP.cpp:785); kept here for reporting (== rows of allres.csv minus 1) */
    return ret;
}

/* generic loop (SEG.cpp ADMM_bqp): runs max_iters iterations from lpo_generic_ineq_init state */
LPO_API int lpo_generic_ineq_run(lpo_solver *s) {
    int iter;
    update_expression(s); /* SEG.cpp:1462-1539 builds the same matrices before the loop (idempotent at iter 0) */
    for (iter = 0; iter < s->max_iters; ++iter) {
        int st = ineq_iteration(s, iter, -1, 0, NULL);
        if (st == 1 || st == 2) break;
    }
    s->iter = iter;
    return 1;
}

/* ADMM_lp_iters_l2f, LP.cpp:1098-1574 */
LPO_API int lpo_lp_iters_l2f(lpo_solver *s, int iter_start, int iter_end, const double *vec, int fix_num) {
    int ret = 0;
    int n = s->n, m = s->m;
    /* :1113 x_iters = Zero(n-fix_num, 500) */
    free(s->x_iters);
    s->xit_rows = n - fix_num; s->xit_cols = 500;
    s->x_iters = (double *)calloc((size_t)(s->xit_rows > 0 ? s->xit_rows : 1) * 500, sizeof(double));
    int cc = 0;
    if (fix_num != 0) {
        s->fix_sum += fix_num;
        int *fix_idx = (int *)malloc(sizeof(int) * (size_t)n), *non_fix_idx = (int *)malloc(sizeof(int) * (size_t)n);
        double *fix_val = (double *)malloc(sizeof(double) * (size_t)n);
        int *newid = (int *)malloc(sizeof(int) * (size_t)n);
        int j = 0, k = 0;
        for (int i = 0; i < n; ++i) {                              /* :1135-1162 */
            if (vec[i] == 1) { fix_idx[j] = i; fix_val[j] = 1; newid[i] = -(j + 1); j++; }
            else if (vec[i] == 0) { fix_idx[j] = i; fix_val[j] = 0; newid[i] = -(j + 1); j++; }
            else { non_fix_idx[k] = i; newid[i] = k; k++; }
        }
        /* E1 (kept columns, order preserved) / E2 (fixed columns) :1135-1183 */
        csx_t E1, E2;
        int nnz = s->Ec.ptr[n];
        csx_alloc(&E1, k, m, nnz); csx_alloc(&E2, j, m, nnz);
        int p1 = 0, p2 = 0, c1 = 0, c2 = 0;
        for (int i = 0; i < n; ++i) {
            if (newid[i] >= 0) {
                for (int q = s->Ec.ptr[i]; q < s->Ec.ptr[i + 1]; ++q) { E1.idx[p1] = s->Ec.idx[q]; E1.val[p1] = s->Ec.val[q]; p1++; }
                E1.ptr[++c1] = p1;
            } else {
                for (int q = s->Ec.ptr[i]; q < s->Ec.ptr[i + 1]; ++q) { E2.idx[p2] = s->Ec.idx[q]; E2.val[p2] = s->Ec.val[q]; p2++; }
                E2.ptr[++c2] = p2;
            }
        }
        /* :1192-1206 index bookkeeping */
        int *new_left = (int *)malloc(sizeof(int) * (size_t)(k > 0 ? k : 1));
        s->ret_idx = (int *)realloc(s->ret_idx, sizeof(int) * (size_t)(s->n_ret + j + 1));
        s->ret_val = (double *)realloc(s->ret_val, sizeof(double) * (size_t)(s->n_ret + j + 1));
        for (int q = 0; q < j; ++q) { s->ret_idx[s->n_ret + q] = s->left_idx[fix_idx[q]]; s->ret_val[s->n_ret + q] = fix_val[q]; }
        s->n_ret += j;
        for (int q = 0; q < k; ++q) new_left[q] = s->left_idx[non_fix_idx[q]];
        free(s->left_idx); s->left_idx = new_left; s->n_left_idx = k;
        if (n - fix_num == 0) {                                    /* :1212-1217 */
            ret = 1; s->n = 0; iter_end = iter_start;
            csx_free(&E1); csx_free(&E2);
        } else {
            /* :1222-1231 gathers (Eigen evaluates x(non_fix_idx) into a temporary first: no aliasing hazard) */
            double **gv[] = {&s->x, &s->y1, &s->y2, &s->z1, &s->z2};
            for (size_t g = 0; g < 5; ++g) { double *v = *gv[g]; for (int q = 0; q < k; ++q) v[q] = v[non_fix_idx[q]]; }
            if (lpo_norm(s->x, k) < 1e-3) ret = 1;                 /* :1223 */
            double *b1 = (double *)malloc(sizeof(double) * (size_t)k), *b2 = (double *)malloc(sizeof(double) * (size_t)(j > 0 ? j : 1));
            for (int q = 0; q < k; ++q) b1[q] = s->b[non_fix_idx[q]];
            for (int q = 0; q < j; ++q) b2[q] = s->b[fix_idx[q]];
            s->fix_obj = lpo_dot(b2, fix_val, j);                  /* :1237 compute_cost_lp(x2,b2) = b2.dot(x2) */
            s->prev_sum = s->sum_fix_obj;
            s->sum_fix_obj += s->fix_obj;                          /* :1248 */
            s->prev_obj = s->cur_obj;
            /* :1276-1278  f1 = f - E2 x2  (row-sequential, ascending fixed-column order) */
            {
                csx_t E2r; csx_transpose(&E2r, &E2);
                spmv_rows(&E2r, E2r.val, fix_val, s->t_m0);
                for (int i = 0; i < m; ++i) s->f[i] = s->f[i] - s->t_m0[i];
                csx_free(&E2r);
            }
            s->n = n - fix_num;                                    /* :1295 */
            csx_free(&s->Ec); csx_free(&s->Er);
            s->Ec = E1; s->Ec.nouter = k;
            csx_transpose(&s->Er, &s->Ec);
            memcpy(s->b, b1, sizeof(double) * (size_t)k);
            free(b1); free(b2); csx_free(&E2);
            update_expression(s);                                  /* :1329 */
        }
        free(fix_idx); free(non_fix_idx); free(fix_val); free(newid);
    }
    n = s->n;
    int iter;
    for (iter = iter_start; iter < iter_end; ++iter) {
        int st = ineq_iteration(s, iter, iter_start, 1, &cc);
        if (st == 3) { s->iter = iter; return 1; }                 /* :1450-1454 */
        if (st == 1 || st == 2) { ret = 1; break; }                /* :1505, :1542 */
    }
    s->iter = iter;                                                /* member `iter` is the loop variable (:1341) */
    return ret;
}

/* getters -------------------------------------------------------------------------------------- */
LPO_API int lpo_get_n(const lpo_solver *s) { return s->n; }
LPO_API int lpo_get_m(const lpo_solver *s) { return s->m; }
LPO_API int lpo_get_org_n(const lpo_solver *s) { return s->org_n; }
LPO_API int lpo_get_iter(const lpo_solver *s) { return s->iter; }
LPO_API long lpo_get_cg_iters(const lpo_solver *s) { return s->cg_iters_total; }
LPO_API long lpo_get_admm_iters(const lpo_solver *s) { return s->admm_iters_total; }
LPO_API double lpo_get_cur_bin_obj(const lpo_solver *s) { return s->cur_obj; }                      /* LP.cpp:1644 */
LPO_API double lpo_cal_obj(const lpo_solver *s) { return s->n != 0 ? s->sum_fix_obj + s->cur_obj : s->sum_fix_obj; } /* :1630-1642 */
LPO_API double lpo_get_sum_fix_obj(const lpo_solver *s) { return s->sum_fix_obj; }
LPO_API void lpo_get_scalars(const lpo_solver *s, double *out /*8*/) {
    out[0] = s->rho1; out[1] = s->rho2; out[2] = s->rho4; out[3] = s->gamma; out[4] = s->std_obj; out[5] = s->cur_obj;
    out[6] = s->best_bin_obj; out[7] = (double)s->obj_len;
}
/* get_final_x_sol LP.cpp:1668-1685: the relaxed x of the current (compacted) problem */
LPO_API void lpo_get_final_x_sol(const lpo_solver *s, double *out) { memcpy(out, s->x, sizeof(double) * (size_t)s->n); }
LPO_API void lpo_get_state(const lpo_solver *s, double *x, double *y1, double *y2, double *z1, double *z2, double *y3, double *z4) {
    size_t nb = sizeof(double) * (size_t)s->n, mb = sizeof(double) * (size_t)s->m;
    if (x) memcpy(x, s->x, nb); if (y1) memcpy(y1, s->y1, nb); if (y2) memcpy(y2, s->y2, nb);
    if (z1) memcpy(z1, s->z1, nb); if (z2) memcpy(z2, s->z2, nb);
    if (y3) memcpy(y3, s->y3, mb); if (z4) memcpy(z4, s->z4, mb);
}
/* get_x_sol LP.cpp:1648-1665: binary solution in original indexing (length org_n); entries never written stay as
 * given in `out` (the reference returns uninitialised memory there) */
LPO_API void lpo_get_x_sol(const lpo_solver *s, double *out) {
    for (int q = 0; q < s->n_ret; ++q) out[s->ret_idx[q]] = s->ret_val[q];
    if (s->n != 0)
        for (int q = 0; q < s->n_left_idx; ++q) out[s->left_idx[q]] = (s->x[q] >= 0.5) ? 1.0 : 0.0;
}
/* get_x_iters_d LP.cpp:1616-1627: row-major (rows x ws) copy of the first ws columns */
LPO_API int lpo_get_x_iters(const lpo_solver *s, int ws, double *out) {
    for (int i = 0; i < s->xit_rows; ++i)
        for (int j = 0; j < ws; ++j) out[(size_t)i * ws + j] = s->x_iters[(size_t)j * (size_t)s->xit_rows + i];
    return s->xit_rows;
}
LPO_API int lpo_get_x_iters_rows(const lpo_solver *s) { return s->xit_rows; }
/* check_infeasible_lpbox LP.cpp:1577-1591 (current E, relaxed x) */
LPO_API int lpo_check_infeasible_lpbox(lpo_solver *s) {
    int inf = 0;
    spmv_rows(&s->Er, s->Er.val, s->x, s->t_m0);
    for (int i = 0; i < s->m; ++i) if (!(s->t_m0[i] <= 1.0)) inf++;
    return inf;
}
/* check_infeasible_l2f LP.cpp:1593-1612 (original E, assembled binary x) */
LPO_API int lpo_check_infeasible_l2f(lpo_solver *s) {
    int inf = 0;
    double *xs = (double *)calloc((size_t)(s->org_n > 0 ? s->org_n : 1), sizeof(double));
    lpo_get_x_sol(s, xs);
    csx_t Eor; csx_transpose(&Eor, &s->Eorg);
    double *t = (double *)calloc((size_t)(Eor.nouter > 0 ? Eor.nouter : 1), sizeof(double));
    spmv_rows(&Eor, Eor.val, xs, t);
    for (int i = 0; i < Eor.nouter; ++i) if (!(t[i] <= 1.0)) inf++;
    free(t); free(xs); csx_free(&Eor);
    return inf;
}

/* single PCG solve on an explicit row-compressed matrix (SEG.cpp:272-342) -- unit-test entry */
LPO_API int lpo_pcg_csr(int n, const int *rowptr, const int *colidx, const double *val, const double *rhs, double *x,
                        const double *invdiag, double tol, int maxit) {
    double *r = (double *)malloc(sizeof(double) * (size_t)n), *p = (double *)malloc(sizeof(double) * (size_t)n),
           *tmp = (double *)malloc(sizeof(double) * (size_t)n);
    int it = 0;
    lpo_spmv_csr(n, rowptr, colidx, val, x, tmp);
    for (int i = 0; i < n; ++i) r[i] = rhs[i] - tmp[i];
    double rn = lpo_sqnorm(rhs, n);
    if (rn == 0) { memset(x, 0, sizeof(double) * (size_t)n); goto done; }
    {
        double thr = tol * tol * rn; if (!(thr > DBL_MIN)) thr = DBL_MIN;
        double r2 = lpo_sqnorm(r, n);
        if (r2 < thr) goto done;
        for (int i = 0; i < n; ++i) p[i] = invdiag[i] * r[i];
        double absNew = lpo_dot(r, p, n);
        while (it < maxit) {
            lpo_spmv_csr(n, rowptr, colidx, val, p, tmp);
            double alpha = absNew / lpo_dot(p, tmp, n);
            for (int k = 0; k < n; ++k) x[k] = x[k] + alpha * p[k];
            for (int k = 0; k < n; ++k) r[k] = r[k] - alpha * tmp[k];
            r2 = lpo_sqnorm(r, n);
            if (r2 < thr) { it++; break; }
            for (int k = 0; k < n; ++k) tmp[k] = invdiag[k] * r[k];
            double absOld = absNew;
            absNew = lpo_dot(r, tmp, n);
            double beta = absNew / absOld;
            for (int k = 0; k < n; ++k) p[k] = tmp[k] + beta * p[k];
            it++;
        }
    }
done:
    free(r); free(p); free(tmp);
    return it;
}
