/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's UNCONSTRAINED Lp-Box ADMM path
 * (graph-cut image segmentation):   min x'Ax + b'x,  x in {0,1}^n.
 *
 * SEG.cpp = Segmentation/Segmentation/cython/src/LPboxADMMsolver.cpp.  Same arithmetic rules as lpbox_oracle.c
 * (SURVEY.md §8c): fp64, no FMA, row-sequential SpMV, Eigen SSE2 reduction order, expression association as written.
 * Parity status: PINNED against the reference binary's `ADMM_bqp_unconstrained` (tests/test_seg_oracle.py) and its
 * exported graph-builder helpers.
 */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define LPO_API __attribute__((visibility("default")))

double lpo_sum(const double *v, long n);
double lpo_dot(const double *a, const double *b, long n);
double lpo_sqnorm(const double *a, long n);
double lpo_norm(const double *a, long n);
void lpo_spmv_csr(int nrows, const int *rowptr, const int *colidx, const double *val, const double *x, double *y);
int lpo_pcg_csr(int n, const int *rowptr, const int *colidx, const double *val, const double *rhs, double *x,
                const double *invdiag, double tol, int maxit);

typedef struct {
    /* hyper-parameters (SEG.cpp:659-672) */
    double stop_threshold, std_threshold;
    int max_iters;
    double initial_rho;
    int rho_change_step;
    double gamma_val, learning_fact, history_size, gamma_factor, pcg_tol;
    int pcg_maxiters;
    /* problem: A row-compressed with explicit diagonal, current (compacted) and original */
    int n, org_n;
    int *rp, *ci; double *av;          /* A_ptr */
    int *orp, *oci; double *oav;       /* org_A */
    double *b, *ob;
    double cconst;
    /* temp_mat = 2A + (rho1+rho2) I : same pattern as A, own values; position of each diagonal entry */
    double *mv; int *dpos;
    double *x, *y1, *y2, *z1, *z2, *invd, *t0, *t1, *t2;
    double rho1, rho2, prho1, prho2, gamma, ratio;
    int rhoUpdated;
    double *obj_list; long obj_len, obj_cap;
    double std_obj, cur_obj, best_bin_obj;
    int *left_idx; int *ret_idx; double *ret_val; int n_ret;
    double *x_iters; int xit_rows;
    long cg_iters_total, admm_iters_total;
    int iter;
} sego_solver;

LPO_API sego_solver *sego_create(void) {
    sego_solver *s = (sego_solver *)calloc(1, sizeof(sego_solver));
    s->std_obj = 1.0; s->rhoUpdated = 1;
    return s;
}
LPO_API void sego_destroy(sego_solver *s) {
    if (!s) return;
    free(s->rp); free(s->ci); free(s->av); free(s->orp); free(s->oci); free(s->oav); free(s->b); free(s->ob);
    free(s->mv); free(s->dpos); free(s->x); free(s->y1); free(s->y2); free(s->z1); free(s->z2); free(s->invd);
    free(s->t0); free(s->t1); free(s->t2); free(s->obj_list); free(s->left_idx); free(s->ret_idx); free(s->ret_val); free(s->x_iters);
    free(s);
}
LPO_API void sego_params_seg(sego_solver *s) {   /* SEG.cpp:659-672 */
    s->std_threshold = 1e-6; s->gamma_val = 1.0; s->gamma_factor = 0.99; s->initial_rho = 5; s->learning_fact = 1 + 3.0 / 100;
    s->history_size = 5; s->rho_change_step = 5; s->stop_threshold = 1e-3; s->max_iters = (int)1e4; s->pcg_tol = 1e-3;
    s->pcg_maxiters = (int)1e3;
}
LPO_API void sego_set_params(sego_solver *s, double stop_threshold, double std_threshold, int max_iters, double initial_rho,
                             int rho_change_step, double gamma_val, double learning_fact, double history_size,
                             double gamma_factor, double pcg_tol, int pcg_maxiters) {
    s->stop_threshold = stop_threshold; s->std_threshold = std_threshold; s->max_iters = max_iters; s->initial_rho = initial_rho;
    s->rho_change_step = rho_change_step; s->gamma_val = gamma_val; s->learning_fact = learning_fact; s->history_size = history_size;
    s->gamma_factor = gamma_factor; s->pcg_tol = pcg_tol; s->pcg_maxiters = pcg_maxiters;
}

static int *idup(const int *p, size_t n) { int *q = (int *)malloc(sizeof(int) * (n ? n : 1)); memcpy(q, p, sizeof(int) * n); return q; }
static double *ddup(const double *p, size_t n) { double *q = (double *)malloc(sizeof(double) * (n ? n : 1)); memcpy(q, p, sizeof(double) * n); return q; }

/* A: row-compressed, ascending columns, every row must store its diagonal (the graph builder keeps explicit zeros) */
LPO_API int sego_set_problem(sego_solver *s, int n, const int *rowptr, const int *colidx, const double *val, const double *b, double c) {
    int nnz = rowptr[n];
    for (int i = 0; i < n; ++i) {
        int found = 0;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            if (k > rowptr[i] && colidx[k] <= colidx[k - 1]) return -1;
            if (colidx[k] == i) found = 1;
        }
        if (!found) return -2;
    }
    s->n = s->org_n = n;
    s->rp = idup(rowptr, (size_t)n + 1); s->ci = idup(colidx, (size_t)nnz); s->av = ddup(val, (size_t)nnz);
    s->orp = idup(rowptr, (size_t)n + 1); s->oci = idup(colidx, (size_t)nnz); s->oav = ddup(val, (size_t)nnz);
    s->b = ddup(b, (size_t)n); s->ob = ddup(b, (size_t)n);
    s->cconst = c;
    return 0;
}

/* temp_mat = 2 * A; temp_mat.diagonal() += rho1 + rho2   (SEG.cpp:784-786, :1054-1057) */
static void build_temp_mat(sego_solver *s) {
    int n = s->n, nnz = s->rp[n];
    free(s->mv); free(s->dpos);
    s->mv = (double *)malloc(sizeof(double) * (size_t)(nnz ? nnz : 1));
    s->dpos = (int *)malloc(sizeof(int) * (size_t)(n ? n : 1));
    for (int k = 0; k < nnz; ++k) s->mv[k] = 2 * s->av[k];
    for (int i = 0; i < n; ++i)
        for (int k = s->rp[i]; k < s->rp[i + 1]; ++k)
            if (s->ci[k] == i) { s->dpos[i] = k; s->mv[k] += s->rho1 + s->rho2; }
}

/* compute_cost(x, A, b) = x.dot(A x) + b.dot(x)   (SEG.cpp:568-572) */
static double compute_cost(sego_solver *s, const double *x) {
    lpo_spmv_csr(s->n, s->rp, s->ci, s->av, x, s->t2);
    double val = lpo_dot(x, s->t2, s->n);
    double val2 = lpo_dot(s->b, x, s->n);
    return val + val2;
}

static double std_obj_of(const double *list, long sz, int history) {   /* SEG.cpp std_dev + compute_std_obj */
    long begin = (sz <= history) ? 0 : sz - history, end = sz, size = end - begin;
    double mean = 0, sd = 0;
    for (long i = begin; i < end; ++i) mean += list[i];
    mean /= (double)size;
    for (long i = 0; i < size; ++i) sd += (list[begin + i] - mean) * (list[begin + i] - mean);
    sd /= (double)(size - 1);
    double r = (sd == 0) ? 0.0 : pow(sd, 1.0 / 2);
    return r / fabs(list[sz - 1]);
}

/* ADMM_bqp_unconstrained_init minus the image front-end (SEG.cpp:747-810); x0 == NULL -> zeros (:761-762) */
LPO_API int sego_init(sego_solver *s, const double *x0, int use_seg_params) {
    if (use_seg_params) sego_params_seg(s);
    int n = s->n;
    size_t nb = sizeof(double) * (size_t)(n ? n : 1);
    double **vs[] = {&s->x, &s->y1, &s->y2, &s->z1, &s->z2, &s->invd, &s->t0, &s->t1, &s->t2};
    for (size_t i = 0; i < sizeof(vs) / sizeof(vs[0]); ++i) { free(*vs[i]); *vs[i] = (double *)calloc(1, nb); }
    if (x0) memcpy(s->x, x0, sizeof(double) * (size_t)n);
    memcpy(s->y1, s->x, sizeof(double) * (size_t)n); memcpy(s->y2, s->x, sizeof(double) * (size_t)n);
    s->rho1 = s->rho2 = s->prho1 = s->prho2 = s->initial_rho;
    s->gamma = s->gamma_val;
    build_temp_mat(s);
    s->best_bin_obj = compute_cost(s, s->x);
    free(s->left_idx); s->left_idx = (int *)malloc(sizeof(int) * (size_t)(n ? n : 1));
    for (int i = 0; i < n; ++i) s->left_idx[i] = i;
    free(s->ret_idx); free(s->ret_val); s->ret_idx = NULL; s->ret_val = NULL; s->n_ret = 0;
    s->obj_len = 0; s->std_obj = 1.0; s->cur_obj = 0; s->rhoUpdated = 1; s->iter = 0;
    s->cg_iters_total = 0; s->admm_iters_total = 0;
    return 1;
}

/* one iteration of the unconstrained loop (SEG.cpp:1222-1357 == :1094-1188).  0 continue, 1 y-stop, 2 std-stop */
static int unc_iteration(sego_solver *s, int iter, int record, int *cc) {
    int n = s->n;
    for (int i = 0; i < n; ++i) {                                   /* y1 :1223-1229 */
        double t = s->x[i] + s->z1[i] / s->rho1;
        s->y1[i] = (t > 1) ? 1.0 : ((t < 0) ? 0.0 : t);
    }
    for (int i = 0; i < n; ++i) s->y2[i] = (s->x[i] + s->z2[i] / s->rho2) - 0.5;   /* y2 :1231-1234 */
    {
        double nrm = lpo_norm(s->y2, n);
        double c = pow((double)n, 1.0 / 2);
        for (int i = 0; i < n; ++i) s->y2[i] = s->y2[i] * c / (2 * nrm) + 0.5;
    }
    if (iter != 0 && s->rhoUpdated) {                               /* :1240-1243 */
        double d = (s->prho1 + s->prho2) * s->ratio;
        for (int i = 0; i < n; ++i) s->mv[s->dpos[i]] += d;
    }
    for (int i = 0; i < n; ++i) s->t0[i] = s->rho1 * s->y1[i] + s->rho2 * s->y2[i] - (s->b[i] + s->z1[i] + s->z2[i]);   /* :1246 */
    if (s->rhoUpdated) {                                            /* :1252-1255 */
        for (int i = 0; i < n; ++i) { double d = s->mv[s->dpos[i]]; s->invd[i] = (d != 0.0) ? 1.0 / d : 1.0; }
        s->rhoUpdated = 0;
    }
    memcpy(s->x, s->y1, sizeof(double) * (size_t)n);                /* :1257 */
    int cg = lpo_pcg_csr(n, s->rp, s->ci, s->mv, s->t0, s->x, s->invd, s->pcg_tol, s->pcg_maxiters);   /* :1261 */
    s->cg_iters_total += cg; s->admm_iters_total += 1;
    if (record && cc) {                                             /* :1131-1134 */
        if (*cc < 10) for (int i = 0; i < n; ++i) s->x_iters[(size_t)(*cc) * (size_t)s->xit_rows + i] = s->x[i];
        (*cc)++;
    }
    {
        double g1 = s->gamma * s->rho1, g2 = s->gamma * s->rho2;    /* :1280-1281 */
        for (int i = 0; i < n; ++i) s->z1[i] = s->z1[i] + g1 * (s->x[i] - s->y1[i]);
        for (int i = 0; i < n; ++i) s->z2[i] = s->z2[i] + g2 * (s->x[i] - s->y2[i]);
    }
    {
        double temp0 = lpo_norm(s->x, n);                           /* :1285-1292 */
        if (!(temp0 > 2.2204e-16)) temp0 = 2.2204e-16;
        for (int i = 0; i < n; ++i) s->t0[i] = s->x[i] - s->y1[i];
        double c1 = lpo_norm(s->t0, n) / temp0;
        for (int i = 0; i < n; ++i) s->t0[i] = s->x[i] - s->y2[i];
        double c2 = lpo_norm(s->t0, n) / temp0;
        if (c1 <= s->stop_threshold && c2 <= s->stop_threshold) return 1;
    }
    if ((iter + 1) % s->rho_change_step == 0) {                     /* :1295-1303 */
        s->prho1 = s->rho1; s->prho2 = s->rho2;
        s->rho1 = s->learning_fact * s->rho1; s->rho2 = s->learning_fact * s->rho2;
        double g = s->gamma * s->gamma_factor;
        s->gamma = (g < 1.0) ? 1.0 : g;
        s->rhoUpdated = 1; s->ratio = s->learning_fact - 1.0;
    }
    {
        double obj = compute_cost(s, s->x);                         /* :1308-1319 */
        if (s->obj_len == s->obj_cap) { s->obj_cap = s->obj_cap ? 2 * s->obj_cap : 1024; s->obj_list = (double *)realloc(s->obj_list, sizeof(double) * (size_t)s->obj_cap); }
        s->obj_list[s->obj_len++] = obj;
        if ((double)s->obj_len >= s->history_size) s->std_obj = std_obj_of(s->obj_list, s->obj_len, (int)s->history_size);
        if (s->std_obj <= s->std_threshold) return 2;
    }
    for (int i = 0; i < n; ++i) s->t1[i] = (s->x[i] >= 0.5) ? 1.0 : 0.0;   /* :1323-1331 */
    s->cur_obj = compute_cost(s, s->t1);
    if (s->best_bin_obj >= s->cur_obj) s->best_bin_obj = s->cur_obj;
    return 0;
}

/* ADMM_bqp_unconstrained_legacy (SEG.cpp:1200-1380): returns int(cur_obj + _c) */
LPO_API int sego_legacy(sego_solver *s) {
    int iter;
    for (iter = 0; iter < s->max_iters; ++iter) {
        int st = unc_iteration(s, iter, 0, NULL);
        if (st) break;
    }
    s->iter = iter;
    for (int i = 0; i < s->n; ++i) s->t1[i] = (s->x[i] >= 0.5) ? 1.0 : 0.0;   /* :1366-1367 */
    s->cur_obj = compute_cost(s, s->t1);
    return (int)(s->cur_obj + s->cconst);
}

/* ADMM_bqp_unconstrained_l2f (SEG.cpp:917-1195) */
LPO_API int sego_l2f(sego_solver *s, int iter_start, int iter_end, const double *vec, int fix_num) {
    int ret = 0, n = s->n;
    free(s->x_iters);
    s->xit_rows = n - fix_num;                                      /* :924 x_iters = Zero(n-fix_num, 10) */
    s->x_iters = (double *)calloc((size_t)(s->xit_rows > 0 ? s->xit_rows : 1) * 10, sizeof(double));
    int cc = 0;
    if (fix_num != 0) {
        int *newid = (int *)malloc(sizeof(int) * (size_t)n), *fix_idx = (int *)malloc(sizeof(int) * (size_t)n), *non = (int *)malloc(sizeof(int) * (size_t)n);
        double *fval = (double *)malloc(sizeof(double) * (size_t)n);
        int j = 0, k = 0;
        for (int i = 0; i < n; ++i) {                               /* :944-972 */
            if (vec[i] == 1) { fix_idx[j] = i; fval[j] = 1; newid[i] = j; j++; }
            else if (vec[i] == 0) { fix_idx[j] = i; fval[j] = 0; newid[i] = j; j++; }
            else { non[k] = i; newid[i] = k; k++; }
        }
        /* Ma = A[keep,keep], Mb = A[keep,fix]  (:973-1015): rows in kept order, columns renumbered, ascending */
        int nnz = s->rp[n];
        int *arp = (int *)calloc((size_t)k + 1, sizeof(int)), *aci = (int *)malloc(sizeof(int) * (size_t)(nnz ? nnz : 1));
        double *aav = (double *)malloc(sizeof(double) * (size_t)(nnz ? nnz : 1));
        double *mbx = (double *)calloc((size_t)(k ? k : 1), sizeof(double));
        int q = 0;
        for (int r = 0; r < k; ++r) {
            int i = non[r];
            double acc = 0.0;                                       /* (Mb x2)_r, row-sequential ascending fixed column */
            for (int e = s->rp[i]; e < s->rp[i + 1]; ++e) {
                int c = s->ci[e];
                if (vec[c] == -1) { aci[q] = newid[c]; aav[q] = s->av[e]; q++; }
                else acc = acc + s->av[e] * fval[newid[c]];
            }
            arp[r + 1] = q; mbx[r] = acc;
        }
        /* bookkeeping (:1017-1026) */
        s->ret_idx = (int *)realloc(s->ret_idx, sizeof(int) * (size_t)(s->n_ret + j + 1));
        s->ret_val = (double *)realloc(s->ret_val, sizeof(double) * (size_t)(s->n_ret + j + 1));
        for (int t = 0; t < j; ++t) { s->ret_idx[s->n_ret + t] = s->left_idx[fix_idx[t]]; s->ret_val[s->n_ret + t] = fval[t]; }
        s->n_ret += j;
        int *nl = (int *)malloc(sizeof(int) * (size_t)(k ? k : 1));
        for (int t = 0; t < k; ++t) nl[t] = s->left_idx[non[t]];
        free(s->left_idx); s->left_idx = nl;
        if (n - fix_num == 0) {                                     /* :1028-1032 */
            ret = 1; s->n = 0; iter_end = iter_start;
            free(arp); free(aci); free(aav);
        } else {
            double **gv[] = {&s->x, &s->y1, &s->y2, &s->z1, &s->z2};   /* :1035-1041 */
            for (size_t g = 0; g < 5; ++g) { double *v = *gv[g]; for (int t = 0; t < k; ++t) v[t] = v[non[t]]; }
            for (int t = 0; t < k; ++t) s->b[t] = 2 * mbx[t] + s->b[non[t]];   /* :1051-1052  b = 2 Mb x2 + b1 */
            free(s->rp); free(s->ci); free(s->av);
            s->rp = arp; s->ci = aci; s->av = aav;
            s->n = k;
            build_temp_mat(s);                                      /* :1054-1057 at the CURRENT rho */
        }
        free(mbx); free(newid); free(fix_idx); free(non); free(fval);
    }
    int iter;
    for (iter = iter_start; iter < iter_end; ++iter) {
        int st = unc_iteration(s, iter, 1, &cc);
        if (st) { ret = 1; break; }                                 /* :1148-1153, :1180-1185 */
    }
    s->iter = iter;
    return ret;
}

LPO_API int sego_get_n(const sego_solver *s) { return s->n; }
LPO_API int sego_get_org_n(const sego_solver *s) { return s->org_n; }
LPO_API int sego_get_iter(const sego_solver *s) { return s->iter; }
LPO_API long sego_get_cg_iters(const sego_solver *s) { return s->cg_iters_total; }
LPO_API long sego_get_admm_iters(const sego_solver *s) { return s->admm_iters_total; }
LPO_API double sego_get_cur_obj(const sego_solver *s) { return s->cur_obj; }
LPO_API void sego_get_state(const sego_solver *s, double *x, double *y1, double *y2, double *z1, double *z2) {
    size_t nb = sizeof(double) * (size_t)s->n;
    if (x) memcpy(x, s->x, nb); if (y1) memcpy(y1, s->y1, nb); if (y2) memcpy(y2, s->y2, nb);
    if (z1) memcpy(z1, s->z1, nb); if (z2) memcpy(z2, s->z2, nb);
}
/* get_x_sol (SEG.cpp:895-915) */
LPO_API void sego_get_x_sol(const sego_solver *s, double *out) {
    for (int q = 0; q < s->n_ret; ++q) out[s->ret_idx[q]] = s->ret_val[q];
    if (s->n != 0) for (int q = 0; q < s->n; ++q) out[s->left_idx[q]] = (s->x[q] >= 0.5) ? 1.0 : 0.0;
}
/* get_final_obj (SEG.cpp:868-893): cost of the assembled binary x on the ORIGINAL problem + _c */
LPO_API double sego_get_final_obj(const sego_solver *s) {
    int n = s->org_n;
    double *xx = (double *)calloc((size_t)(n ? n : 1), sizeof(double)), *t = (double *)calloc((size_t)(n ? n : 1), sizeof(double));
    sego_get_x_sol(s, xx);
    lpo_spmv_csr(n, s->orp, s->oci, s->oav, xx, t);
    double obj = lpo_dot(xx, t, n) + lpo_dot(s->ob, xx, n);
    free(xx); free(t);
    return obj + s->cconst;
}
LPO_API int sego_get_x_iters(const sego_solver *s, int ws, double *out) {   /* get_x_iters_d: row-major (rows x ws) */
    for (int i = 0; i < s->xit_rows; ++i)
        for (int j = 0; j < ws; ++j) out[(size_t)i * ws + j] = (j < 10) ? s->x_iters[(size_t)j * (size_t)s->xit_rows + i] : 0.0;
    return s->xit_rows;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Graph builder (SEG.cpp:55-81, :144-248, :727-758): grey image (nr x nc, row-major uint8) -> A (row-compressed,
 * <= 7 stored entries per row incl. explicit zeros), b, c.  Outputs: rowptr[n+1], colidx/val[7n] (caller allocates).
 * ---------------------------------------------------------------------------------------------------------------- */
LPO_API int sego_build_graph(const unsigned char *img, int nr, int nc, int *rowptr, int *colidx, double *val, double *b, double *c_out) {
    const int n = nr * nc;
    double *I = (double *)malloc(sizeof(double) * (size_t)n);          /* I(r, c) = grey / 263  (:727), kept row-major here */
    for (int k = 0; k < n; ++k) I[k] = (double)img[k] / 263.0;
    double *v = (double *)malloc(sizeof(double) * (size_t)n);          /* vectorize(): column-major flattening (:46-53) */
    for (int cc = 0; cc < nc; ++cc) for (int r = 0; r < nr; ++r) v[(size_t)cc * nr + r] = I[(size_t)r * nc + cc];
    /* unary costs (:55-81, :734-743), sigma = 0.1, b = 0.6, f1 = f2 = 0.2, then rounded (std::round) */
    const double sigma = 0.1, bb = 0.6, f1 = 0.2, f2 = 0.2;
    const double cst = log(2.0 * M_PI) / 2.0 + log(sigma);
    double csum = 0.0;
    double *u1 = (double *)malloc(sizeof(double) * (size_t)n);
    for (int k = 0; k < n; ++k) {
        double ab = pow(v[k] - bb, 2.0) / (2 * sigma * sigma) + cst;
        double aa = exp(-pow(v[k] - f1, 2.0) / (2 * sigma * sigma)) + exp(-pow(v[k] - f2, 2) / (2 * sigma * sigma));
        double af = -log(aa + DBL_EPSILON) + cst + log(2.0);
        double U1 = round(ab), U2 = round(af);
        b[k] = U2 - U1;                                              /* :230 */
        u1[k] = U1;
    }
    csum = lpo_sum(u1, n);                                           /* c = U1.sum()  (:238) */
    free(u1);
    /* pairwise weights (:173-224): sigma_img = sample std of v (NOT squared) */
    double mean = lpo_sum(v, n) / (double)n;
    double *d2 = (double *)malloc(sizeof(double) * (size_t)n);
    for (int k = 0; k < n; ++k) d2[k] = (v[k] - mean) * (v[k] - mean);
    double sig = sqrt(lpo_sum(d2, n) / (double)(n - 1));
    free(d2);
    /* pairs over linear row-major index p = i*nc + j with offsets (a,b), a != b (:144-171); intensities looked up as
     * image(p % nr, p / nr) i.e. through the COLUMN-major flattening (the reference's index mismatch, :192-193) */
    int q = 0;
    for (int i = 0; i < nr; ++i)
        for (int j = 0; j < nc; ++j) {
            const int p = i * nc + j;
            rowptr[p] = q;
            double wsum = 0.0;
            int dq = -1;
            /* ascending column order: (a,b) = (-1,0), (-1,1), (0,-1), [diag], (0,1), (1,-1), (1,0) */
            const int oa[7] = {-1, -1, 0, 0, 0, 1, 1}, ob[7] = {0, 1, -1, 0, 1, -1, 0};
            for (int t = 0; t < 7; ++t) {
                const int a = oa[t], bo = ob[t];
                if (a == 0 && bo == 0) { dq = q; colidx[q] = p; val[q] = 0.0; q++; continue; }
                if (i + a < 0 || i + a >= nr || j + bo < 0 || j + bo >= nc) continue;
                const int p2 = (i + a) * nc + (j + bo);
                const double i1 = I[(size_t)(p % nr) * nc + (p / nr)], i2 = I[(size_t)(p2 % nr) * nc + (p2 / nr)];
                double w = round(3 * exp(-(pow(i1 - i2, 2.0) / sig)));
                colidx[q] = p2; val[q] = -w; q++;                    /* A = -W (:232) */
            }
            /* We = -A * ones (row-sequential over the stored row, diagonal contributes 0) ; A.diag += We (:234-237) */
            for (int e = rowptr[p]; e < q; ++e) wsum = wsum + (-val[e]) * 1.0;
            val[dq] = val[dq] + wsum;
            /* A = 2*A (:238) then A_ptr = _A/2 (:756): exact for these small integers */
        }
    rowptr[n] = q;
    *c_out = csum;
    free(I); free(v);
    return q;
}
