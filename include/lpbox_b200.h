/*
 * lpbox_b200 -- C ABI of the B200-native Lp-Box ADMM solver (drop-in for the reference's solver objects).
 *
 * Reference aliases (all under the upstream tree):
 *   LP.cpp / LP.h / LP.pxd / LP.pyx = LinerProgramming/LinearProgramming/cython_solver/{LPboxADMMsolver.cpp,.h,.pxd,lpbox.pyx}
 *   SEG.cpp / SEG.pxd / SEG.pyx     = Segmentation/Segmentation/cython/src/{LPboxADMMsolver.cpp,.pxd,lpbox.pyx}
 *
 * Every entry point below replaces one method the reference's Cython layer binds (LP.pxd:4-21, SEG.pxd:4-17);
 * the method it replaces is cited next to it.  Plain pointers and sizes only; all buffers are HOST memory owned by
 * the caller unless the name says `_dev`.  Functions return >= 0 on success (the reference's own return value where
 * it has one) and a negative LPBOX_E_* code on error.  There is no CPU fallback: if no CUDA device is usable the
 * create functions fail with LPBOX_E_CUDA.
 *
 * Threading: a handle is stateful and not re-entrant (like the reference object); different handles may be used
 * from different threads.  One handle = one CUDA stream.
 */
#ifndef LPBOX_B200_H
#define LPBOX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LPBOX_E_INVALID      (-1)  /* bad argument / wrong call order */
#define LPBOX_E_CUDA         (-2)  /* CUDA runtime error (see lpbox_last_error) */
#define LPBOX_E_UNSUPPORTED  (-3)  /* problem too large for the on-chip kernels */
#define LPBOX_E_IO           (-4)  /* file could not be read */

typedef struct lpbox_batch lpbox_batch;   /* B independent instances resident in HBM */

/* hyper-parameters of the iteration (LP.h:115-146).  lpbox_params_lp() = LP.cpp:491-507,
 * lpbox_params_seg() = SEG.cpp:659-672. */
typedef struct lpbox_params {
    double stop_threshold;
    double std_threshold;
    int    max_iters;
    double initial_rho;
    int    rho_change_step;
    double gamma_val;
    double learning_fact;
    double history_size;
    double projection_lp;
    double gamma_factor;
    double pcg_tol;
    int    pcg_maxiters;
} lpbox_params;

void lpbox_params_lp(lpbox_params *p);
void lpbox_params_seg(lpbox_params *p);

/* per-instance result row ("iteration log out"; the reference writes file_idx,-obj,iters,seconds to allres.csv,
 * LP.cpp:1081) */
typedef struct lpbox_log_row {
    int32_t iters;        /* ADMM iterations executed */
    int32_t status;       /* 0 = ran to iter_end, 1 = ||x-y1||,||x-y2|| stop, 2 = objective-std stop, 3 = CG alpha<0, 4 = all variables fixed */
    int64_t cg_iters;     /* total CG iterations */
    double  obj;          /* cal_obj(): sum_fix_obj + cur_obj (minimised, i.e. -revenue)  LP.cpp:1630-1642 */
    double  cur_bin_obj;  /* get_curBinObj()  LP.cpp:1644 */
    int32_t n_left;       /* variables still free */
    int32_t infeasible;   /* violated rows of E_orig x <= 1 for the assembled binary x (check_infeasible_l2f, LP.cpp:1593-1612) */
} lpbox_log_row;

const char *lpbox_last_error(void);
int  lpbox_device_count(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Batched LP (inequality-constrained) solver:   min b'x  s.t.  E x <= f,  x in {0,1}^n        (configs 1, 2, 5)
 * ------------------------------------------------------------------------------------------------------------- */

/* Creates B instances on `device`.  Instance i has m[i] rows, n[i] columns; its matrix is column-compressed
 * (what readFile/readSparseMat build, LP.cpp:2416-2444): colptr_all holds the B colptr arrays back to back
 * (n[i]+1 entries each, each starting at 0), rowidx_all / val_all the concatenated row indices (strictly ascending
 * inside a column) / values.  val_all == NULL means all stored entries are 1.0 (auction instances,
 * generate_instances.py:359).  b_all: concatenated objective vectors as the SOLVER sees them (readFile negates the
 * bid prices, LP.cpp:2520).  f_all == NULL means f = 1 (LP.cpp:2522).  hist_cap = number of iterates per window
 * kept for get_x_iters (the reference keeps 500 columns, LP.cpp:1113); 0 disables the history.
 * Replaces: LPboxADMMsolver(int) + readFile  (LP.pxd:7,9). */
lpbox_batch *lpbox_batch_create(int device, int B, const int32_t *m, const int32_t *n, const int32_t *colptr_all,
                                const int32_t *rowidx_all, const double *val_all, const double *b_all,
                                const double *f_all, int hist_cap);
void lpbox_batch_destroy(lpbox_batch *h);

/* Overrides the hyper-parameters that lpbox_batch_init() would set (header setters LP.h:511-575).
 * variant bit 0: keep the plain loop's `iter != iter_start` guard + z4 assignment (LP.cpp:934,:920-921);
 * variant bit 1: keep the CG alpha<0 bail-out (LP.cpp:307).  Default after create: both set (LP behaviour);
 * 0 reproduces the generic ADMM_bqp loop of SEG.cpp:1590-1760. */
int lpbox_batch_set_params(lpbox_batch *h, const lpbox_params *p, int variant);

/* ADMM_lp_iters_init (LP.cpp:489-763) for every instance; x0_all == NULL means x = 1 (LP.cpp:583-586). */
int lpbox_batch_init(lpbox_batch *h, const double *x0_all);

/* print_fix_info == 2 (LP.cpp:903-909): the plain loop records every iterate too (at most hist_cap per call); read with
 * lpbox_batch_get_x_iters */
int lpbox_batch_set_record_history(lpbox_batch *h, int on);
/* ADMM_lp_iters(iter_start, iter_end) (LP.cpp:766-1095) for every instance that has not stopped.
 * ret[i] (may be NULL) receives the reference's return value (1 only for the objective-std stop). */
int lpbox_batch_iters(lpbox_batch *h, int iter_start, int iter_end, int32_t *ret);

/* ADMM_lp_iters_l2f(iter_start, iter_end, vec, num) (LP.cpp:1098-1574).  vec_all: concatenated fix vectors over the
 * CURRENT variables of each instance (n_cur[i] entries in {1,0,-1}); num[i] = number of entries != -1 (0 = no fix;
 * vec of that instance is then ignored, as in the reference).  vec_all may be NULL if every num[i] is 0.
 * Instances whose previous call returned 1 are skipped (the reference driver stops calling, LP.trainer:521). */
int lpbox_batch_iters_l2f(lpbox_batch *h, int iter_start, int iter_end, const double *vec_all, const int32_t *num,
                          int32_t *ret);

/* Arithmetic mode of the window kernel.  0 (default) = PARITY: every fp64 operation in the reference's order, iterates
 * bit-identical to the reference's compiled Eigen code (LP.cpp:251-335, :766-1095).  1 = FAST: the same iteration with tree
 * reductions over all warps and fused multiply-adds in the vector updates; iterates agree with the reference to rounding
 * level per step only, so final solutions can differ -- opt-in, never used for parity claims. */
int lpbox_batch_set_mode(lpbox_batch *h, int mode);

/* Plain Lp-Box ADMM to convergence for every instance: update_expression + ADMM_lp_iters(0, max_iters) (LP.cpp:766-1095),
 * one launch; log (B rows, may be NULL) as lpbox_batch_results fills it. */
int lpbox_batch_solve(lpbox_batch *h, int max_iters, lpbox_log_row *log /* B rows, may be NULL */);

/* ---- device-resident window loop: window -> policy -> threshold -> compact without host copies of the iterates ----
 * (LP.trainer:510-535).  The policy network itself runs outside this library on `scores`/`input` DEVICE buffers
 * (PyTorch or the bf16 policy kernel); everything is ordered on one stream. */
/* use the caller's CUDA stream (cudaStream_t) for all work of this handle, e.g. torch's current stream */
int lpbox_batch_set_stream(lpbox_batch *h, void *cuda_stream);
/* device pointer of the iterate history: instance i at element offset sum_{k<i} hist_cap*n[k], layout [iteration][n[i]] fp64 */
void *lpbox_batch_hist_dev(lpbox_batch *h);
/* Packs the last window's history of every still-active instance into out_dev as fp32 rows [row][ws] -- the reference's
 * `xiters.reshape(n_left, 20, ws/20).astype(float32)` (LP.trainer:524-530) -- rows of the active instances back to
 * back in instance order.  Returns the number of rows (call with out_dev == NULL to size the buffer). */
int64_t lpbox_batch_policy_input_dev(lpbox_batch *h, int ws, float *out_dev, int64_t capacity_rows);
/* deter_fix_2 (LP.trainer:101-135) on device scores laid out like the rows above: p > hi -> fix 1, p < lo -> fix 0,
 * else keep; instances with <= min_fix fixes get none (LP.trainer:533-535).  The result feeds the next _dev window. */
int lpbox_batch_apply_scores_dev(lpbox_batch *h, const float *scores_dev, double hi, double lo, int min_fix);
/* ADMM_lp_iters_l2f for all active instances with the fix vectors produced by lpbox_batch_apply_scores_dev (none if it
 * was not called since the last window).  Returns the number of instances still active afterwards. */
int lpbox_batch_iters_l2f_dev(lpbox_batch *h, int iter_start, int iter_end);

/* The whole early-fixing loop of the reference driver (LP.trainer:510-545) in one call: windows of `ws` iterations
 * (ADMM_lp_iters_l2f), after each window the policy scores every free variable of every running instance from the window's
 * iterate history (the reshape of LP.trainer:524-530), deter_fix_2 (LP.trainer:101-135: p > hi -> 1, p < lo -> 0; instances
 * with <= min_fix fixes get none, :533-535) and the compaction of the next window's prologue -- all on the device, on the
 * handle's stream.  The active list and the policy-input row offsets are built on the device; the host reads 24 bytes per
 * window.  `policy`: lpbox_policy_create (declared below).  Create the batch with hist_cap >= ws; ws must be a multiple of
 * rho_change_step.  log / x_bits as lpbox_batch_results (either may be NULL).
 * Replaces the Python loop around solve_iter_l2f / get_x_iters_2d / deter_fix_2 (LP.trainer:510-545). */
typedef struct lpbox_policy lpbox_policy;
typedef struct {
    int32_t windows;        /* windows run */
    int64_t policy_rows;    /* variable-windows scored by the policy */
    double  device_ms;      /* CUDA-event time of the whole loop on the handle's stream */
} lpbox_l2f_stats;
int lpbox_batch_solve_l2f(lpbox_batch *h, lpbox_policy *policy, int ws, int max_iter, double hi, double lo, int min_fix,
                          lpbox_log_row *log, uint8_t *x_bits, int row_stride_bytes, lpbox_l2f_stats *stats);

/* Optional extension, OFF by default (= the reference, whose deter_fix_2 trusts the policy): when on, a fix-to-one proposal is
 * only applied if every constraint row it touches still has capacity for it under the current right-hand side and it is the
 * best-scored proposal of that row, so the fixed part of the solution satisfies E x <= f by construction; rejected proposals
 * stay free.  Applies to lpbox_batch_apply_scores_dev and lpbox_batch_solve_l2f. */
int lpbox_batch_set_fix_guard(lpbox_batch *h, int on);

/* getters; `i` = instance index ---------------------------------------------------------------------------------- */
int lpbox_batch_size(const lpbox_batch *h);
int lpbox_batch_get_n(lpbox_batch *h, int i);            /* get_n()      LP.h:392 */
int lpbox_batch_get_m(lpbox_batch *h, int i);
int lpbox_batch_get_org_n(lpbox_batch *h, int i);
int lpbox_batch_get_iter(lpbox_batch *h, int i);         /* get_iter()   LP.h:357 */
double lpbox_batch_cal_obj(lpbox_batch *h, int i);       /* cal_obj()    LP.cpp:1630-1642 */
double lpbox_batch_get_cur_bin_obj(lpbox_batch *h, int i); /* get_curBinObj() LP.cpp:1644 */
/* get_x_sol(): binary solution in original indexing, out has org_n entries; entries the reference leaves
 * uninitialised keep the caller's value  (LP.cpp:1648-1665) */
int lpbox_batch_get_x_sol(lpbox_batch *h, int i, double *out);
/* get_final_x_sol(): relaxed x of the current (compacted) problem, n entries  (LP.cpp:1668-1685) */
int lpbox_batch_get_final_x_sol(lpbox_batch *h, int i, double *out);
/* left_idx (LP.h:246): original ids of the variables still free, n entries; returns n */
int lpbox_batch_get_left_idx(lpbox_batch *h, int i, int32_t *out);
/* get_x_iters_d(ws): row-major (n_cur x ws) window history  (LP.cpp:1616-1627).  Returns n_cur. */
int lpbox_batch_get_x_iters(lpbox_batch *h, int i, int ws, double *out);
int lpbox_batch_check_infeasible_lpbox(lpbox_batch *h, int i);  /* LP.cpp:1577-1591 */
int lpbox_batch_check_infeasible_l2f(lpbox_batch *h, int i);    /* LP.cpp:1593-1612 */
/* full iterate state of instance i (any pointer may be NULL): x,y1,y2,z1,z2 (n), y3,z4 (m) -- parity sweeps */
int lpbox_batch_get_state(lpbox_batch *h, int i, double *x, double *y1, double *y2, double *z1, double *z2,
                          double *y3, double *z4);
/* one log row per instance, plus packed binary solutions (bit j of byte j/8 of row i; row stride = (max org_n+7)/8)
 * -- the "final gather" payload of SURVEY.md §8e.  Either pointer may be NULL. */
int lpbox_batch_results(lpbox_batch *h, lpbox_log_row *log, uint8_t *x_bits, int row_stride_bytes);
/* device time (ms, CUDA events on the handle's stream) and number of kernel launches of the last solve/iters call */
double lpbox_batch_last_kernel_ms(const lpbox_batch *h);
int64_t lpbox_batch_launch_count(const lpbox_batch *h);
/* launch configuration of the window kernel: out4 = {grid (persistent CTAs), dynamic shared memory bytes per CTA,
 * threads per CTA, shared memory bytes of the early-fix kernel} */
int lpbox_batch_config(const lpbox_batch *h, int32_t *out4);
/* bytes copied host->device / device->host by this handle so far (counted from the buffers actually copied) */
int64_t lpbox_batch_h2d_bytes(const lpbox_batch *h);
int64_t lpbox_batch_d2h_bytes(const lpbox_batch *h);

/* ---------------------------------------------------------------------------------------------------------------
 * Batched UNCONSTRAINED solver (graph-cut image segmentation):   min x'Ax + b'x,  x in {0,1}^n            (config 3)
 * Replaces the Segmentation experiment's LPboxADMMsolver (SEG.pxd:4-17).  Problems do not fit on chip; one CTA streams
 * one image (see csrc/seg_kernels.cuh).
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct lpbox_seg_batch lpbox_seg_batch;

/* B problems given as row-compressed symmetric A (columns strictly ascending inside a row, EVERY row stores its
 * diagonal entry -- explicit zeros allowed, as the reference's graph builder produces, SEG.cpp:213-219), b and the
 * constant c (`_c`, SEG.cpp:238).  rowptr_all: the B rowptr arrays back to back (n[i]+1 entries each, starting at 0). */
lpbox_seg_batch *lpbox_seg_create_csr(int device, int B, const int32_t *n, const int32_t *rowptr_all, const int32_t *colidx_all,
                                      const double *val_all, const double *b_all, const double *c, int hist_cap);
/* B grey images (row-major uint8, nr[i] x nc[i], already scaled to the node budget): uploads the PIXELS only and runs the
 * reference's graph construction (get_unary_cost / get_binary_cost / get_A_b_from_cost, SEG.cpp:55-81,144-248,727-758) on
 * the device, one CTA per image, straight into the solver's CSR buffers.  Replaces the image front-end
 * of ADMM_bqp_unconstrained_init (SEG.cpp:705-758) minus imread/resize. */
lpbox_seg_batch *lpbox_seg_create_images(int device, int B, const uint8_t *pixels_all, const int32_t *nr, const int32_t *nc,
                                         int hist_cap);
void lpbox_seg_destroy(lpbox_seg_batch *h);
/* the graph the batch holds for image / problem i (what the device builder produced): rowptr[n+1], colidx/val[nnz], b[n], *c;
 * returns nnz */
int lpbox_seg_get_graph(lpbox_seg_batch *h, int i, int32_t *rowptr, int32_t *colidx, double *val, double *b, double *c);
/* the HOST graph builder alone (single image; the literal restatement the device builder is tested against): outputs
 * rowptr[n+1], colidx/val[<= 7n], b[n], *c; returns nnz */
int lpbox_seg_build_graph(const uint8_t *pixels, int nr, int nc, int32_t *rowptr, int32_t *colidx, double *val, double *b,
                          double *c_out);
int lpbox_seg_set_params(lpbox_seg_batch *h, const lpbox_params *p);        /* default: lpbox_params_seg() */
/* ADMM_bqp_unconstrained_init (SEG.cpp:658-810) state part; x0_all == NULL means x = 0 (SEG.cpp:761-762) */
int lpbox_seg_init(lpbox_seg_batch *h, const double *x0_all);
/* ADMM_bqp_unconstrained_legacy (SEG.cpp:1200-1380) for every image; energy[i] = int(cur_obj + _c) as it returns */
int lpbox_seg_solve(lpbox_seg_batch *h, int32_t *energy);
/* ADMM_bqp_unconstrained_l2f(iter_start, iter_end, vec, num) (SEG.cpp:917-1195), argument conventions as
 * lpbox_batch_iters_l2f; the compaction A <- A[keep,keep], b <- 2 A[keep,fix] x_fix + b[keep] runs on the device */
int lpbox_seg_iters_l2f(lpbox_seg_batch *h, int iter_start, int iter_end, const double *vec_all, const int32_t *num, int32_t *ret);
int lpbox_seg_get_x_iters(lpbox_seg_batch *h, int i, int ws, double *out);   /* get_x_iters_d(ws)  SEG.cpp:833-845 */
int lpbox_seg_size(const lpbox_seg_batch *h);
int lpbox_seg_get_n(lpbox_seg_batch *h, int i);                             /* get_n()      */
int lpbox_seg_get_org_n(lpbox_seg_batch *h, int i);                         /* get_org_n()  */
int lpbox_seg_get_iter(lpbox_seg_batch *h, int i);
int lpbox_seg_get_x_sol(lpbox_seg_batch *h, int i, double *out);            /* get_x_sol()      SEG.cpp:895-915 */
double lpbox_seg_get_final_obj(lpbox_seg_batch *h, int i);                  /* get_final_obj()  SEG.cpp:868-893 */
int lpbox_seg_get_state(lpbox_seg_batch *h, int i, double *x, double *y1, double *y2, double *z1, double *z2);
int lpbox_seg_results(lpbox_seg_batch *h, lpbox_log_row *log);
double lpbox_seg_last_kernel_ms(const lpbox_seg_batch *h);
int64_t lpbox_seg_launch_count(const lpbox_seg_batch *h);
int64_t lpbox_seg_h2d_bytes(const lpbox_seg_batch *h);
int64_t lpbox_seg_d2h_bytes(const lpbox_seg_batch *h);

/* ---------------------------------------------------------------------------------------------------------------
 * Sparse adversarial attack: Lp-Box ADMM on the pixel mask G, fp32, batched over images                     (config 4)
 * Replaces the tensor arithmetic of update_G / loop / update_G_l2f (SparseAttack/main_ori.py:626-743, :502-623,
 * :376-499); the attacked classifier stays in PyTorch.  ALL pointers are DEVICE pointers (torch CUDA tensors, fp32,
 * [n_img][n_elem] with n_elem = channels*H*W); `stream` is a cudaStream_t.  Segments (the reference's SLIC blocks B,
 * main_ori.py:147-158) must partition the elements: seg_of[e] = segment of element e, seg_ptr/seg_elems = elements
 * grouped by segment (nseg+1 / n_elem entries), shared by all images (seg_per_image = 0) or one set per image.
 * ------------------------------------------------------------------------------------------------------------- */
/* steps 1-2 of an iteration (main_ori.py:652-664) + the classifier input (:670-672):
 * y1 = clamp(G + z1/rho1, 0, 1); y2 = sqrt(n)/2 * s/||s|| + 1/2, s = G + z2/rho2 - 1/2 (utils.py:8-16);
 * y3 = group-lasso prox of C = G + z3/rho3; image_s = (clamp(images + G*eps, minpix, maxpix) - mean) / std */
int lpbox_sa_pre_dev(void *stream, int n_img, int n_elem, int n_chan, int nseg, int seg_per_image, const float *G, const float *z1,
                     const float *z2, const float *z3, const float *images, const float *eps, const int32_t *seg_ptr,
                     const int32_t *seg_elems, const int32_t *seg_of, const float *mean, const float *stdv, double rho1, double rho2,
                     double rho3, double lambda2, double minpix, double maxpix, float *y1, float *y2, float *y3, float *image_s);
/* steps 3-4 (main_ori.py:697-721) given grad_in = dLoss/d image_s from the classifier: chain rule to G, grad_G, gradient
 * step, z1..z4 updates; writes the new G into hist_slot as well when it is not NULL (G_iters, main_ori.py:586) */
int lpbox_sa_post_dev(void *stream, int n_img, int n_elem, int n_chan, float *G, float *z1, float *z2, float *z3, float *z4,
                      const float *y1, const float *y2, const float *y3, const float *grad_in, const float *images, const float *eps,
                      const float *nw, const float *stdv, double lambda1, const float *lambda1_img /* [n_img] or NULL */, double rho1,
                      double rho2, double rho3, double rho4, double step, double k, double minpix, double maxpix, float *hist_slot);
/* The perturbation step either side of update_G (update_epsilon, main_ori.py:310-354):
 * image_s = (clamp(images + eps*G, minpix, maxpix) - mean) / std  (:317-319) */
int lpbox_sa_eps_pre_dev(void *stream, int n_img, int n_elem, int n_chan, const float *images, const float *eps, const float *G,
                         const float *mean, const float *stdv, double minpix, double maxpix, float *image_s);
/* eps <- eps - step * (2*eps*G*G*w*w + lambda1 * dLoss/d eps) (:341-343) given grad_in = dLoss/d image_s */
int lpbox_sa_eps_post_dev(void *stream, int n_img, int n_elem, int n_chan, float *eps, const float *G, const float *grad_in,
                          const float *images, const float *nw, const float *stdv, double lambda1, const float *lambda1_img,
                          double step, double minpix, double maxpix);
/* compute_statistics (utils.py:77-96) per image: out9[img] = {G_sum, L0, L1, L2, Li, WL1, WL2, WLi, ||G*eps*w||_2^2} */
int lpbox_sa_stats_dev(void *stream, int n_img, int n_elem, const float *images, const float *eps, const float *G, const float *nw,
                       double minpix, double maxpix, float *out9);
/* update_G_l2f's rewrite of G from policy scores (main_ori.py:476-485): p > hi -> 1, p < lo -> 0, else `last`;
 * counts2[0], counts2[1] receive the number of ones / zeros fixed */
int lpbox_sa_apply_policy_dev(void *stream, int64_t n, const float *scores, const float *last, double hi, double lo, float *G,
                              int32_t *counts2);

/* ---------------------------------------------------------------------------------------------------------------
 * Early-fixing policy network on the tensor cores (GraphAttentionEncoder / MLPEncoder forward, LP.mha:202-304, eval
 * mode): bf16 tcgen05 GEMMs with fp32 TMEM accumulators + small CUDA-core kernels (csrc/policy_kernels.cu).
 * `packed`: fp32 host buffer in the layout written by lpbox/policy_kernel.py:pack_policy.  tokens <= 32.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct lpbox_policy lpbox_policy;
lpbox_policy *lpbox_policy_create(int device, int tokens, int n_layers, const float *packed, int64_t n_packed, int64_t chunk_rows);
void lpbox_policy_destroy(lpbox_policy *p);
/* scores_dev[r] = sigmoid(net(input_dev[r])), input_dev: DEVICE fp32 [rows][tokens*5] (= the packed window history of
 * lpbox_batch_policy_input_dev), scores_dev: DEVICE fp32 [rows]; ordered on `stream` (cudaStream_t) */
int lpbox_policy_forward_dev(lpbox_policy *p, void *stream, const float *input_dev, int64_t rows, float *scores_dev);
int64_t lpbox_policy_launch_count(const lpbox_policy *p);
/* the tcgen05 GEMM alone (tests): C[M][N] = A[M][K] . W[N][K]^T (+ bias[n]) (ReLU); bf16 DEVICE row-major; N % 128 == 0, K % 64 == 0 */
int lpbox_gemm_bf16_dev(void *stream, const void *A, const void *W, void *C, int64_t M, int N, int K, const float *bias, int relu);
/* the fused multi-head-attention sublayer alone (tests): out = (X + Wo MHA(X)) * scale + shift (LP.mha:58-122 + skip + eval-mode
 * BatchNorm); X, out: bf16 DEVICE [M][128] with M = variables * T rows, T = 20, 10 or 5 tokens per variable; Wqkv: bf16 [384][128]
 * (q | k | v, head-major inside each third); Wo: bf16 [128][128] */
int lpbox_mha_fused_dev(void *stream, const void *X, const void *Wqkv, const void *Wo, const float *scale, const float *shift,
                        void *out, int64_t M, int T);
/* the fused feed-forward sublayer alone (tests): out = (X + W2 relu(W1 X + b1) + b2) * scale + shift  (LP.mha:140-160 with the
 * eval-mode BatchNorm folded into scale / shift); X, out: bf16 DEVICE [M][128]; W1: bf16 [512][128]; W2: bf16 [128][512] */
int lpbox_ff_fused_dev(void *stream, const void *X, const void *W1, const float *b1, const void *W2, const float *b2,
                       const float *scale, const float *shift, void *out, int64_t M);

/* ---------------------------------------------------------------------------------------------------------------
 * File format of the reference: data/instance/<k>_<j>/instance_<i>_{C,b}.txt under `root`
 * (readFile, LP.cpp:2446-2545).  Arrays are malloc()ed by the library; release with lpbox_free().
 * ------------------------------------------------------------------------------------------------------------- */
int lpbox_read_instance(const char *root, int i, int k, int j, int32_t *m, int32_t *n, int32_t **colptr,
                        int32_t **rowidx, double **val, double **b);
void lpbox_free(void *p);

/* Synthetic instances: `count` auctions from the "arbitrary" scheme of Leyton-Brown et al. (EC-00 §4.3) with the
 * parameterisation of the reference generator (generate_instances.py:137-140; add_item_prob as passed at :396 = 0.7).
 * Own random stream (not numpy's): same distribution, different instances.  Outputs are malloc()ed, concatenated in
 * the layout lpbox_batch_create takes (colptr: count x (n_bids+1); price: count x n_bids, POSITIVE bid prices --
 * negate for b); release with lpbox_free().  threads <= 0: all host cores. */
/* Diagnostic, host only (no device needed): shared-memory wavefronts of the operand gathers of one E v and one E^T w of an instance
 * under the window kernel's slot assignment (mode 0: slots by descending stored length; 1: the bank-aware assignment the library uses).
 * out[4] = {E v actual, E v conflict-free, E^T w actual, E^T w conflict-free}; cap = 512 / 1024 / 2048 (slots of the kernel shape).
 * No counterpart in the reference (it has no GPU path); used by tests/ and tools/ to track the quality of the assignment. */
int lpbox_debug_gather_wavefronts(int m, int n, const int32_t *colptr, const int32_t *rowidx, int cap, int mode, int64_t *out);

int lpbox_gen_auctions(uint64_t seed, int count, int n_items, int n_bids, double add_item_prob, int threads,
                       int32_t **m_out, int32_t **colptr_out, int32_t **rowidx_out, double **price_out);

#ifdef __cplusplus
}
#endif
#endif /* LPBOX_B200_H */
