// Shared host/device data layout of the batched LP (inequality) Lp-Box ADMM solver.
//
// Everything an instance owns lives in HBM between kernel launches (this is the reference's "solver state lives in
// the C++ object between solve_iter_l2f calls", LP.h:199-262); a window kernel pulls one instance on chip, runs
// its iterations entirely in registers / shared memory and writes the state back.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define LPB_HD __host__ __device__ __forceinline__
#else
#define LPB_HD inline
#endif

namespace lpb {

// Hyper-parameters (LP.h:115-146) + loop-variant switches.
struct Params {
    double stop_threshold, std_threshold;
    int max_iters;
    double initial_rho;
    int rho_change_step;
    double gamma_val, learning_fact;
    int history_size;        // compared as `obj_list.size() >= history_size` (double in the reference; integral values only)
    double gamma_factor, pcg_tol;
    int pcg_maxiters;
    int guard_first_iter;    // LP.cpp:934 + :920-921 (plain LP loop)
    int alpha_bailout;       // LP.cpp:307
};

enum Status : int { RUNNING = 0, STOP_Y = 1, STOP_STD = 2, STOP_CG = 3, STOP_EMPTY = 4 };

// Per-instance scalars.  One struct per instance in HBM.
struct InstState {
    int n0, m0, nnz0;        // capacities (original problem)
    int rcap, ccap;          // capacities of the sliced-ELL index arrays (32-entry groups)
    int n, m, nnz;           // current (after early fixing) -- get_n()
    int iter;                // loop variable `iter` as left by the last window (get_iter())
    int status;              // Status of the last window
    int done;                // != 0: the reference driver would have stopped calling (ret == 1)
    int last_ret;            // return value of the last ADMM_lp_iters / _l2f call
    int rhoUpdated;          // LP.h:208
    int unit;                // all stored values of E are 1.0
    int n_ret;               // length of ret_idx_prev / ret_val_prev
    int fix_sum;
    int xit_rows, xit_cols;  // shape of x_iters of the last l2f window (LP.cpp:1113), cols actually recorded
    int norm_small;          // LP.cpp:1223 flag (x_sol.norm() < 1e-3 after a fix)
    double rho1, rho2, rho4, prho1, prho2, prho4, gamma, ratio;
    double D;                // every diagonal entry of _2A_plus_rho1_rho2 (identical for all j)
    double r4s;              // unit case: the single value stored in rho4_E_transpose
    double std_obj, cur_obj, best_bin_obj, sum_fix_obj, fix_obj, prev_obj, prev_sum;
    long long obj_len;       // obj_list.size()
    double obj_ring[16];     // last 16 entries of obj_list (history_size <= 16)
    long long cg_iters, admm_iters;
};

// Two images of the sparsity pattern of E per instance, all uint16:
//
// (1) CsrLayout -- both compressed orientations (rowptr/colidx, colptr/rowidx).  Lives in HBM only; used by the
//     set-up / early-fix kernels and the host getters.
// (2) EllLayout -- the image the window kernel stages into shared memory with one 1-D TMA bulk copy: both
//     orientations in PADDED SLICED-ELL form.  "Slot" s is the work item AND the home of row rperm[s] (column
//     cperm[s]): the thread that owns slot s keeps that row's / column's vector entries and computes its sparse
//     product.  Slots are sorted by descending stored length; a slice = 32 consecutive slots, stored column-major and
//     padded to the length of its longest slot: entry k of lane l of slice w sits at idx[ell_pos(32 sptr[w], l, k, W)] -- the
//     entries of a lane are stored in PAIRS (k, k + 1), pair-column-major, so that one 32-bit load fetches two offsets; an odd
//     last entry forms a final plain column (the slice still takes 32 W entries).
//     An entry is not an index but the SHARED-WINDOW ADDRESS of the operand it gathers (sbase = start of the window
//     kernel's dynamic shared memory): sbase + 8 * (slot of that column) for the row image (operand vector G at offset
//     0), sbase + gather_base(cap) + 8 * (slot of that row) for the column image (operand vector T1); padding entries
//     point at one shared 0.0 (sbase + zero_off(cap)) -- adding +0.0 to a sum that started at +0.0 never changes it, so padding is value-neutral and
//     every lane of a warp runs the same trip count.  Only WHO computes a row / column product and WHERE operands sit
//     changes; the order of operations inside each product (ascending inner index) is the reference's.
//     sptr / idx arrays are staged; the two slot -> index permutations stay in global memory (read once per window).
struct CsrLayout {
    int o_rowptr, o_colptr, o_colidx, o_rowidx, bytes;
};
struct EllLayout {
    int o_rsptr, o_csptr, o_ridx, o_cidx, o_rperm, o_cperm, bytes;   // [0, o_rperm) is staged; o_cidx.. may stay in L2
};
LPB_HD constexpr int a16(int x) { return (x + 15) & ~15; }
LPB_HD CsrLayout csr_layout(int n0, int m0, int nnz0) {
    CsrLayout L;
    L.o_rowptr = 0;
    L.o_colptr = a16(2 * (m0 + 1));
    L.o_colidx = L.o_colptr + a16(2 * (n0 + 1));
    L.o_rowidx = L.o_colidx + a16(2 * nnz0);
    L.bytes = L.o_rowidx + a16(2 * nnz0);
    return L;
}
// rcap / ccap: capacity of the row / column ELL index arrays in 32-entry groups (sum of the slice widths at creation;
// early fixing can only shrink them)
// The SpMV loops prefetch one batch of offsets (4 steps x 32 lanes x 2 bytes = 256 bytes) past the end of a slice without a
// bounds test; what follows the arrays only has to be readable: the row offsets are followed by the column offsets, those by the
// permutations (global memory) / by 256 spare bytes at the end of the window kernel's shared memory.
// position (in entries) of entry k of lane l in a slice of width W that starts at entry `base`
LPB_HD constexpr int ell_pos(int base, int l, int k, int W) {
    return k < (W & ~1) ? base + (k >> 1) * 64 + 2 * l + (k & 1) : base + (W >> 1) * 64 + l;
}
LPB_HD EllLayout ell_layout(int n0, int m0, int rcap, int ccap) {
    EllLayout L;
    const int nsr = (m0 + 31) / 32, nsc = (n0 + 31) / 32;
    L.o_rsptr = 0;
    L.o_csptr = a16(2 * (nsr + 1));
    L.o_ridx = L.o_csptr + a16(2 * (nsc + 1));
    L.o_cidx = L.o_ridx + a16(64 * rcap);
    L.o_rperm = L.o_cidx + a16(64 * ccap);
    L.o_cperm = L.o_rperm + a16(2 * m0);
    L.bytes = L.o_cperm + a16(2 * n0);
    return L;
}
// Chain-major reduction buffers: element j of an n-vector sits at (j & 3) * CH + (j >> 2) (Eigen's four interleaved
// chains become four contiguous runs).  CH = 2 (mod 4) doubles: CH * 8 bytes is then an odd multiple of 16 modulo 128, so the
// four chains start in four different 16-byte bank groups, and two buffers laid out back to back start 64 bytes apart modulo
// 128 -> eight lanes reading 16 bytes each (two reductions side by side) touch every bank exactly once.
LPB_HD constexpr int chain_stride(int np) {
    const int c = np / 4 + 3;                  // terms per chain, the slot of the tail elements, two never-written doubles
    return 4 * ((c - 2 + 3) / 4) + 2;          // (the last two doubles of a buffer stay 0.0: padding operand of the z4 copy)
}
// shared-memory map of the window kernel that the image refers to; cap = T * EPT of the kernel variant in use (compile-time in
// the kernel: runtime region sizes cost registers there)
LPB_HD constexpr int zero_off(int cap) { return 4 * chain_stride(cap) * 8; }     // the shared 0.0 (end of the G region)
LPB_HD constexpr int gather_base(int cap) { return zero_off(cap) + 16; }         // T1 starts here

// Device view of a batch.
struct BatchView {
    int B;
    int hist_cap;
    const long long *off_n;    // [B+1] element offsets of the n-vectors (stride n0 rounded up to 2)
    const long long *off_m;    // [B+1]
    const long long *off_pat;  // [B+1] byte offsets of the sliced-ELL blobs (16-byte aligned)
    const long long *off_csr;  // [B+1] byte offsets of the CSR/CSC blobs
    const long long *off_val;  // [B+1] element offsets of the compressed-order value arrays (nnz0 each); non-unit only
    const long long *off_evr;  // [B+1] element offsets of the row-ELL-order value array (32*rcap each)
    const long long *off_evc;  // [B+1] element offsets of the column-ELL-order value arrays (32*ccap each)
    const long long *off_hist; // [B+1] element offsets of the iterate history (hist_cap * n0 each)
    double *x, *y1, *y2, *z1, *z2, *b, *Pd, *Esq;  // n-vectors
    double *y3, *z4, *f;                           // m-vectors
    unsigned char *pat;                            // sliced-ELL blobs (shared-memory images)
    unsigned char *csr;                            // CSR/CSC blobs
    double *val_r, *val_c;                         // CSR-order / CSC-order values of E (non-unit only)
    double *ev_r, *ev_c, *r4v;                     // ELL-order values: E (row slots), E (column slots), rho4*E^T (column slots)
    InstState *st;
    double *hist;                                  // [cc][n0] per instance (iteration-major, coalesced writes)
    int *left_idx;                                 // [off_n] current -> original variable id
    int *ret_idx;                                  // [off_n] fixed original ids (ret_idx_prev)
    double *ret_val;                               // [off_n]
    const double *pow_tab;                         // pow_tab[k] = pow((double)k, 0.5) from the host libm
    int cap;                                       // T * EPT of the window-kernel variant of this batch
    int sbase;                                     // shared-window address of the window kernel's dynamic shared memory (lp_probe_kernel)
};

struct Launch {
    int iter_start, iter_end;
    int l2f;                 // 1: ADMM_lp_iters_l2f semantics (record history, ret=1 on either stop, CG bail-out returns)
    int record;              // 1: also record the iterate history in the plain loop (print_fix_info == 2, LP.cpp:903-909)
    int skip_done;           // 1: skip instances whose previous call returned "stop" (batch drivers)
    int n_work;              // number of work items
    const int *work;         // instance ids (NULL: identity)
    int *counter;            // atomic work counter (device)
    int np, mp;              // shared-memory vector strides (>= max n0, m0 of the batch; even)
    int pat_bytes;           // shared-memory bytes reserved for the sliced-ELL blob
    int evr_elems, evc_elems; // shared-memory doubles reserved for the ELL-order value arrays (0 when unit)
    int tab_len;             // unit case: entries of the shared 1/diag table (longest column of the batch + 1)
    int slice;               // > 0: sliced queue -- at most `slice` iterations per pop, running instances are re-queued (plain batch solves)
    int *ring;               // [ring_cap] sliced queue: instance ids, -1 = not produced yet; the first n_work entries are pre-filled
    int ring_cap;
    int *tail, *finished;    // sliced queue: next free ring slot; number of instances that are done
    int *sm_rank;            // [#SMs] zeroed per launch: arrival order of the CTAs of one SM (rotates the reduction warp)
    double *park;            // [grid][8][cap] per-CTA parking lot (L2-resident) for vectors that are not touched inside PCG
    int *error;              // device flag: set when the shared-window base differs from BatchView::sbase
    int fast;                // 1: fast mode (tree reductions, FMA) -- NOT bit-identical to the reference
};

}  // namespace lpb
