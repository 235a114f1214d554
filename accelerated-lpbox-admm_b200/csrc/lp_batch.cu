// Host side of the batched LP solver + its C ABI (include/lpbox_b200.h).
// Plain CUDA runtime; no torch types cross this boundary.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#include "../../include/lpbox_b200.h"
#include "lp_fix_kernel.cuh"
#include "lp_kernels.cuh"
#include "lp_policy_glue.cuh"

using namespace lpb;

static thread_local std::string g_err;
static void set_err(const std::string &s) { g_err = s; }
void lpbox_set_error(const std::string &s) { g_err = s; }   // shared with seg_batch.cu
extern "C" const char *lpbox_last_error(void) { return g_err.c_str(); }

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            set_err(std::string(#call) + ": " + cudaGetErrorString(e_));                                \
            return LPBOX_E_CUDA;                                                                        \
        }                                                                                               \
    } while (0)

extern "C" int lpbox_device_count(void) {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) return 0;
    return c;
}

extern "C" void lpbox_params_lp(lpbox_params *p) {  // LP.cpp:491-507
    p->stop_threshold = 1e-4; p->std_threshold = 1e-12; p->max_iters = (int)2e4; p->initial_rho = 25;
    p->rho_change_step = 25; p->gamma_val = 1.6; p->learning_fact = 1 + 1.0 / 100; p->history_size = 10;
    p->projection_lp = 2; p->gamma_factor = 0.95; p->pcg_tol = 1e-3; p->pcg_maxiters = (int)1e3;
}
extern "C" void lpbox_params_seg(lpbox_params *p) {  // SEG.cpp:659-672
    p->stop_threshold = 1e-3; p->std_threshold = 1e-6; p->max_iters = (int)1e4; p->initial_rho = 5;
    p->rho_change_step = 5; p->gamma_val = 1.0; p->learning_fact = 1 + 3.0 / 100; p->history_size = 5;
    p->projection_lp = 2; p->gamma_factor = 0.99; p->pcg_tol = 1e-3; p->pcg_maxiters = (int)1e3;
}

template <typename Tp>
struct DevBuf {
    Tp *p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        n = count;
        return cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(Tp));
    }
    void free_() { if (p) cudaFree(p); p = nullptr; }
};

struct lpbox_batch {
    int device = 0, B = 0, hist_cap = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<int> n0, m0, nnz0;
    std::vector<long long> off_n, off_m, off_pat, off_csr, off_val, off_evr, off_evc, off_hist;
    // host copy of the original problem (check_infeasible_l2f uses org_E_ptr, LP.cpp:1593-1612)
    std::vector<int> h_colptr, h_rowidx;
    std::vector<long long> h_nnz_off, h_cp_off;
    std::vector<double> h_val;
    bool all_unit = true;
    int max_n = 0, max_m = 0, max_nnz = 0, max_pat = 0, max_csr = 0, max_evr = 0, max_evc = 0;
    Params pr{};
    BatchView bv{};
    DevBuf<long long> d_off_n, d_off_m, d_off_pat, d_off_csr, d_off_val, d_off_evr, d_off_evc, d_off_hist, d_off_vec;
    DevBuf<double> d_x, d_y1, d_y2, d_z1, d_z2, d_b, d_Pd, d_Esq, d_y3, d_z4, d_f, d_val_r, d_val_c, d_ev_r, d_ev_c, d_r4v, d_hist,
        d_ret_val, d_pow, d_vec;
    DevBuf<unsigned char> d_pat, d_csr;
    DevBuf<InstState> d_st;
    DevBuf<int> d_left, d_ret_idx, d_counter, d_num;
    std::vector<InstState> h_st;
    bool inited = false;
    int tcfg = 0;  // 0: <128,4>  1: <256,4>  2: <512,4>  3: <256,2> (n <= 512 with twice the threads)  4: <512,2> (n <= 1024)
    int threads = 128, ept = 4;
    int grid = 0;
    size_t smem = 0, fix_smem = 0;
    double last_ms = 0;
    int64_t launches = 0;
    int64_t h2d_bytes = 0, d2h_bytes = 0;
    bool own_stream = true;
    // device-resident window loop
    DevBuf<int> d_active;
    DevBuf<long long> d_row_off;
    std::vector<int> h_active;
    std::vector<long long> h_row_off;
    bool dev_fix_pending = false;
    bool record_plain = false;          // lpbox_batch_set_record_history
    std::vector<int> pat_bytes_i, pat_head_i;   // sliced-ELL image size of every instance, and its size without the column index array
    int pat_smem = 0;                           // bytes of shared memory reserved for the image (< max_pat: the largest images keep their column indices in L2)
    DevBuf<int> d_work;                         // launch order when pat_smem < max_pat (instances with a spilled image first)
    DevBuf<double> d_park;                      // [grid][PK_COUNT][cap] per-CTA parking lot of the window kernel
    DevBuf<int> d_err;                          // [2]: error flag of the window kernel, result of the shared-window probe
    DevBuf<int> d_ring;                         // sliced work queue of plain batch solves: [0] tail, [1] finished, [2..] ring
    int ring_cap = 0;
    DevBuf<float> d_pinp, d_pscore;             // lpbox_batch_solve_l2f: policy input [rows][ws] / scores [rows]
    DevBuf<L2fMeta> d_meta;
    int pinp_ws = 0;
    bool guard = false;                         // lpbox_batch_set_fix_guard: feasibility guard on fix-to-one decisions (not in the reference)
    int cap = 0, nwarps = 0;                    // T * EPT and T / 32 of the window-kernel variant
    int max_col_len = 0, tab_len = 0;           // longest column of the batch; entries of the shared 1/diag table (unit case)
    bool fast = false;                          // lpbox_batch_set_mode: tree reductions + FMA (not bit-identical)
};
static const int SM_RANK_SLOTS = 1024;          // >= highest %smid + 1

// instance-parallel host loops of lpbox_batch_create (sorting, packing): blocks of 64 instances over the host cores
template <typename F>
static void host_parallel_for(int count, F f) {
    int nt = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
    if (count < 512) nt = 1;
    if (nt == 1) { for (int i = 0; i < count; ++i) f(i); return; }
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t)
        pool.emplace_back([&]() {
            for (int i0; (i0 = next.fetch_add(64)) < count;)
                for (int i = i0; i < std::min(i0 + 64, count); ++i) f(i);
        });
    for (auto &th : pool) th.join();
}

// The opt-in limit of a kernel is one value per function and context, NOT per handle: it is always raised to the device maximum, so
// that handles with different shared-memory needs can be alive (and launch from different host threads) at the same time.
template <int T, int EPT, bool UNIT>
static cudaError_t prep_kernel(size_t smem, int smem_optin_max, int *occ) {
    cudaError_t e = cudaFuncSetAttribute(lp_admm_window_kernel<T, EPT, UNIT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin_max);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(lp_admm_window_kernel<T, EPT, UNIT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin_max);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, lp_admm_window_kernel<T, EPT, UNIT, false>, T, smem);
}
template <int T, int EPT, bool UNIT>
static void launch_window(lpbox_batch *h, const Launch &la, int grid) {
    if (h->fast) lp_admm_window_kernel<T, EPT, UNIT, true><<<grid, T, h->smem, h->stream>>>(h->bv, h->pr, la);
    else lp_admm_window_kernel<T, EPT, UNIT, false><<<grid, T, h->smem, h->stream>>>(h->bv, h->pr, la);
}

static int configure(lpbox_batch *h) {
    int dim = std::max(h->max_n, h->max_m);
    const char *force = getenv("LPBOX_TCFG");     // developer override: "256x2" / "512x2" run small shapes with more threads per CTA
    if (dim <= 512) h->tcfg = (force && !strcmp(force, "256x2")) ? 3 : 0;
    else if (dim <= 1024) h->tcfg = (force && !strcmp(force, "512x2")) ? 4 : 1;
    else if (dim <= 2048) h->tcfg = 2;
    else { set_err("max(n, m) > 2048 is not supported by the on-chip kernel"); return LPBOX_E_UNSUPPORTED; }
    if (h->max_nnz > 65535) { set_err("nnz > 65535 is not supported by the on-chip kernel"); return LPBOX_E_UNSUPPORTED; }
    int mp = (h->max_m + 1) & ~1, np = std::max((h->max_n + 1) & ~1, mp);   // m-vectors alias n-sized buffers
    int val_elems = h->all_unit ? 0 : ((h->max_nnz + 1) & ~1);
    static const int cfgT[5] = {128, 256, 512, 256, 512}, cfgE[5] = {4, 4, 4, 2, 2};
    const int T = cfgT[h->tcfg];
    h->threads = T; h->ept = cfgE[h->tcfg];
    h->cap = T * h->ept; h->nwarps = T / 32; h->bv.cap = h->cap;
    h->tab_len = h->all_unit ? h->max_col_len + 1 : 0;
    h->smem = smem_bytes(h->cap, np, mp, h->max_pat, h->all_unit ? 0 : h->max_evr, h->all_unit ? 0 : h->max_evc, h->nwarps, h->tab_len);
    h->fix_smem = fix_smem_bytes((h->max_n + 1) & ~1, mp, h->max_csr, val_elems);
    int dev_smem = 0, sms = 0;
    CK(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    if ((int)h->smem > dev_smem || (int)h->fix_smem > dev_smem) {
        set_err("instance does not fit in shared memory"); return LPBOX_E_UNSUPPORTED;
    }
    bool u = h->all_unit;
    auto occ_at = [&](size_t smem, int *occ) -> cudaError_t {
        switch (h->tcfg) {
            case 0: return u ? prep_kernel<128, 4, true>(smem, dev_smem, occ) : prep_kernel<128, 4, false>(smem, dev_smem, occ);
            case 1: return u ? prep_kernel<256, 4, true>(smem, dev_smem, occ) : prep_kernel<256, 4, false>(smem, dev_smem, occ);
            case 3: return u ? prep_kernel<256, 2, true>(smem, dev_smem, occ) : prep_kernel<256, 2, false>(smem, dev_smem, occ);
            case 4: return u ? prep_kernel<512, 2, true>(smem, dev_smem, occ) : prep_kernel<512, 2, false>(smem, dev_smem, occ);
            default: return u ? prep_kernel<512, 4, true>(smem, dev_smem, occ) : prep_kernel<512, 4, false>(smem, dev_smem, occ);
        }
    };
    int occ = 1, occ_reg = 1;
    CK(occ_at(1024, &occ_reg));          // what the registers allow
    CK(occ_at(h->smem, &occ));
    h->pat_smem = h->max_pat;
    h->d_work.free_();
    int spill_budget = -1;
    // Shared memory is sized for the LARGEST sliced-ELL image of the batch; a few instances with a wide pattern can cost the
    // whole batch one resident CTA per SM.  When that happens the image budget is cut to what full occupancy allows and the
    // (few) instances above it keep their last array -- the column index array -- in global memory (L2) instead.
    const char *forced = getenv("LPBOX_IMAGE_BUDGET");      // tests: force a budget (bytes) to exercise the spilled-image path
    if (u && (occ < occ_reg || forced) && (int)h->pat_bytes_i.size() == h->B) {
        int budget = 0;
        if (forced) {
            budget = atoi(forced) & ~15;
        } else {
            size_t lo = 1024, hi = h->smem;                   // largest dynamic size that still reaches occ_reg
            int o = 0;
            while (hi - lo > 16) {
                size_t mid = ((lo + hi) / 2) & ~(size_t)15;
                CK(occ_at(mid, &o));
                if (o >= occ_reg) lo = mid; else hi = mid;
            }
            const size_t fixed = smem_bytes(h->cap, np, mp, 0, 0, 0, h->nwarps, h->tab_len);
            budget = lo > fixed ? (int)((lo - fixed) & ~(size_t)15) : 0;
        }
        int fit = 0, head_max = 0;
        for (int i = 0; i < h->B; ++i) { fit += h->pat_bytes_i[i] <= budget; head_max = std::max(head_max, h->pat_head_i[i]); }
        if (head_max <= budget && budget < h->max_pat && (forced || fit >= (h->B * 9) / 10)) {
            h->pat_smem = budget;
            h->smem = smem_bytes(h->cap, np, mp, budget, 0, 0, h->nwarps, h->tab_len);
            spill_budget = budget;
            if (getenv("LPBOX_DEBUG")) fprintf(stderr, "[lpbox] image budget %d B (largest %d B): %d of %d instances keep their column indices in L2\n", budget, h->max_pat, h->B - fit, h->B);
        }
        CK(occ_at(h->smem, &occ));
    }
    CK(cudaFuncSetAttribute(lp_fix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dev_smem));
    if (occ < 1) occ = 1;
    h->grid = std::max(1, std::min(h->B, sms * occ));
    // Launch order.  A CTA keeps an instance until it stops, so a launch of more than one wave ends with the instances that were
    // started last: start the ones with the most stored entries first (they tend to run longest: corr(nnz, CG iterations) = 0.5 on
    // the auction batches, tools/dump_iters.py) -- longest-processing-time-first list scheduling with a free predictor.
    // Instances whose image does not fit the shared-memory budget (the widest patterns) lead in any case.
    static const bool no_sort = getenv("LPBOX_NO_SORT") != nullptr;         // experiments
    if (spill_budget >= 0 || (h->B > h->grid && !no_sort && (int)h->nnz0.size() == h->B)) {
        std::vector<int> order(h->B);
        for (int i = 0; i < h->B; ++i) order[i] = i;
        auto spilled = [&](int i) { return spill_budget >= 0 && h->pat_bytes_i[i] > spill_budget; };
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            const bool sa = spilled(a), sb = spilled(b);
            if (sa != sb) return sa;
            return (h->B > h->grid && !no_sort) ? h->nnz0[a] > h->nnz0[b] : false;
        });
        CK(h->d_work.alloc(order.size()));
        CK(cudaMemcpy(h->d_work.p, order.data(), sizeof(int) * order.size(), cudaMemcpyHostToDevice));
    }
    h->d_park.free_();
    CK(h->d_park.alloc((size_t)h->grid * PK_COUNT * (size_t)h->cap));
    // the ELL image stores absolute shared-window addresses: ask where the dynamic shared memory of a kernel without static
    // shared memory starts (the window kernel checks that its own base is the same)
    if (!h->d_err.p) CK(h->d_err.alloc(2));
    CK(cudaMemsetAsync(h->d_err.p, 0, 2 * sizeof(int), h->stream));
    lp_probe_kernel<<<1, 1, 64, h->stream>>>(h->d_err.p + 1);
    CK(cudaGetLastError());
    int probe[2] = {0, 0};
    CK(cudaMemcpyAsync(probe, h->d_err.p, sizeof(probe), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->bv.sbase = probe[1];
    if (h->bv.sbase + gather_base(h->cap) + 8 * (h->cap + 2) > 65535) { set_err("shared-window addresses do not fit in 16 bits"); return LPBOX_E_UNSUPPORTED; }
    return 0;
}

static int run_window(lpbox_batch *h, int iter_start, int iter_end, int l2f, int skip_done, int slice = 0) {
    Launch la{};
    la.iter_start = iter_start; la.iter_end = iter_end; la.l2f = l2f; la.skip_done = skip_done; la.record = h->record_plain ? 1 : 0;
    la.n_work = h->B; la.work = h->d_work.p; la.counter = h->d_counter.p;
    la.mp = (h->max_m + 1) & ~1; la.np = std::max((h->max_n + 1) & ~1, la.mp); la.pat_bytes = h->pat_smem;
    la.evr_elems = h->all_unit ? 0 : h->max_evr; la.evc_elems = h->all_unit ? 0 : h->max_evc; la.tab_len = h->tab_len;
    static const bool no_rotate = getenv("LPBOX_NO_ROTATE") != nullptr;     // experiments: reduction warp fixed to the last warp
    la.sm_rank = no_rotate ? nullptr : h->d_counter.p + 1; la.park = h->d_park.p; la.fast = h->fast ? 1 : 0; la.error = h->d_err.p;
    CK(cudaMemsetAsync(h->d_counter.p, 0, sizeof(int) * (1 + SM_RANK_SLOTS), h->stream));
    const bool slice_force = getenv("LPBOX_SLICE_FORCE") != nullptr;      // tests: slice small batches too
    if (slice > 0 && (h->B >= 2 * h->grid || slice_force) && iter_end > iter_start) {
        // sliced queue: every instance can be re-queued once per slice of its (at most iter_end - iter_start) iterations
        const long long cap = (long long)h->B * ((iter_end - iter_start + slice - 1) / slice + 1);
        if (cap < (1ll << 30)) {
            if (!h->d_ring.p || h->ring_cap < (int)cap) { h->d_ring.free_(); CK(h->d_ring.alloc((size_t)cap + 2)); h->ring_cap = (int)cap; }
            CK(cudaMemsetAsync(h->d_ring.p, 0xff, sizeof(int) * ((size_t)h->ring_cap + 2), h->stream));      // -1 = not produced yet
            std::vector<int> head(h->B + 2);
            head[0] = h->B; head[1] = 0;                                                                     // tail, finished
            for (int i = 0; i < h->B; ++i) head[2 + i] = i;
            if (h->d_work.p) CK(cudaMemcpyAsync(h->d_ring.p + 2, h->d_work.p, sizeof(int) * (size_t)h->B, cudaMemcpyDeviceToDevice, h->stream));
            CK(cudaMemcpyAsync(h->d_ring.p, head.data(), sizeof(int) * (h->d_work.p ? 2 : (size_t)h->B + 2), cudaMemcpyHostToDevice, h->stream));
            CK(cudaStreamSynchronize(h->stream));                                                            // `head` is a local
            la.slice = slice; la.ring = h->d_ring.p + 2; la.ring_cap = h->ring_cap; la.tail = h->d_ring.p; la.finished = h->d_ring.p + 1;
        }
    }
    bool u = h->all_unit;
    switch (h->tcfg) {
        case 0: u ? launch_window<128, 4, true>(h, la, h->grid) : launch_window<128, 4, false>(h, la, h->grid); break;
        case 1: u ? launch_window<256, 4, true>(h, la, h->grid) : launch_window<256, 4, false>(h, la, h->grid); break;
        case 3: u ? launch_window<256, 2, true>(h, la, h->grid) : launch_window<256, 2, false>(h, la, h->grid); break;
        case 4: u ? launch_window<512, 2, true>(h, la, h->grid) : launch_window<512, 2, false>(h, la, h->grid); break;
        default: u ? launch_window<512, 4, true>(h, la, h->grid) : launch_window<512, 4, false>(h, la, h->grid); break;
    }
    CK(cudaGetLastError());
    h->launches += 1;
    return 0;
}

static int sync_states(lpbox_batch *h) {
    h->d2h_bytes += (int64_t)(sizeof(InstState) * (size_t)h->B);
    CK(cudaMemcpyAsync(h->h_st.data(), h->d_st.p, sizeof(InstState) * (size_t)h->B, cudaMemcpyDeviceToHost, h->stream));
    int err = 0;
    CK(cudaMemcpyAsync(&err, h->d_err.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (err) { set_err("window kernel: unexpected shared-window base address (ELL image addresses are invalid)"); return LPBOX_E_CUDA; }
    return 0;
}


// ---- bank-aware slot assignment (host, per instance) -------------------------------------------------------------------------
// The window kernel gathers 8-byte operands from shared memory with addresses that depend on the sparsity pattern; a 64-bit
// access is served half a warp at a time and two lanes of a half-warp collide when their operands sit in the same 8-byte bank
// pair (slot mod 16) at different addresses.  Which slot a row / column occupies is free as far as the arithmetic goes (only
// the order INSIDE each sparse product is the reference's), so it is chosen to avoid collisions:
//   * columns are grouped 16 at a time (= the lanes of one half-warp) in descending length such that the members of a group
//     have distinct chain-major staging banks ((CH (j & 3) + (j >> 2)) mod 16, csrc/lp_types.h chain_stride): the scattered
//     staging stores of the reduction operands become conflict-free;
//   * inside every group of 16 columns (rows) the positions are permuted so that the columns (rows) that one half-warp gathers
//     in one step of E v (E^T w) fall into different bank pairs as far as a greedy assignment manages (one sweep).
// Group membership fixes which operands are gathered together; the position inside the group only moves the bank -- so the two
// permutation problems (columns for E v, rows for E^T w) are independent.
namespace {
struct GatherSets {          // for one orientation: sets of inner items gathered together by one half-warp in one step
    std::vector<int> set_ptr, set_mem;          // members (distinct inner items) of set s: set_mem[set_ptr[s] .. set_ptr[s+1])
    std::vector<int> of_ptr, of_set;            // sets that contain inner item x: of_set[of_ptr[x] .. of_ptr[x+1])
};
// outer items in slot order `order` (groups of 16 consecutive slots gather together), lists[o] = inner items of outer item o ascending
static void build_sets(const std::vector<int> &order, const int *ptr, const int32_t *idx, int n_inner, GatherSets &G) {
    G.set_ptr.assign(1, 0); G.set_mem.clear();
    const int n_outer = (int)order.size();
    std::vector<int> tmp;
    for (int g0 = 0; g0 < n_outer; g0 += 16) {
        const int g1 = std::min(g0 + 16, n_outer);
        int W = 0;
        for (int s = g0; s < g1; ++s) W = std::max(W, ptr[order[s] + 1] - ptr[order[s]]);
        for (int k = 0; k < W; ++k) {
            tmp.clear();
            for (int s = g0; s < g1; ++s) { const int o = order[s]; if (ptr[o] + k < ptr[o + 1]) tmp.push_back(idx[ptr[o] + k]); }
            std::sort(tmp.begin(), tmp.end()); tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
            if (tmp.size() < 2) continue;                                   // a single address can not collide
            G.set_mem.insert(G.set_mem.end(), tmp.begin(), tmp.end());
            G.set_ptr.push_back((int)G.set_mem.size());
        }
    }
    const int ns = (int)G.set_ptr.size() - 1;
    G.of_ptr.assign(n_inner + 1, 0);
    for (int v : G.set_mem) G.of_ptr[v + 1]++;
    for (int x = 0; x < n_inner; ++x) G.of_ptr[x + 1] += G.of_ptr[x];
    G.of_set.resize(G.set_mem.size());
    std::vector<int> fill(G.of_ptr.begin(), G.of_ptr.end() - 1);
    for (int s2 = 0; s2 < ns; ++s2) for (int q = G.set_ptr[s2]; q < G.set_ptr[s2 + 1]; ++q) G.of_set[fill[G.set_mem[q]]++] = s2;
}
// permutes the members of every full group of 16 in `order` (slot order of the INNER items) to lower the bank collisions of G's sets
static void place_in_groups(std::vector<int> &order, const GatherSets &G, int sweeps) {
    const int n = (int)order.size();
    const int ns = (int)G.set_ptr.size() - 1;
    if (ns == 0 || n < 16) return;
    // cnt[s][p]: members of set s placed at position p so far; mx[s] = max_p cnt[s][p] = the wavefronts set s costs.  An item goes to the
    // free position that raises the fewest maxima (the true cost), ties broken by the fewest members already there (pairwise collisions).
    static const bool pairwise_only = getenv("LPBOX_PLACE_PAIRWISE") != nullptr;      // experiments: the round-2a cost function
    std::vector<unsigned char> cnt((size_t)ns * 16, 0), mx(ns, 0);
    std::vector<int> pos(n, -1);                                           // item -> position (0..15) inside its group, -1 = not placed
    auto add = [&](int x, int p) {
        for (int q = G.of_ptr[x]; q < G.of_ptr[x + 1]; ++q) { const int s2 = G.of_set[q]; const unsigned char c = ++cnt[(size_t)s2 * 16 + p]; if (c > mx[s2]) mx[s2] = c; }
    };
    auto remove = [&](int x, int p) {
        for (int q = G.of_ptr[x]; q < G.of_ptr[x + 1]; ++q) {
            const int s2 = G.of_set[q];
            unsigned char *c = &cnt[(size_t)s2 * 16];
            --c[p];
            unsigned char m2 = 0;
            for (int r = 0; r < 16; ++r) m2 = std::max(m2, c[r]);
            mx[s2] = m2;
        }
    };
    for (int sw = 0; sw < sweeps; ++sw)
        for (int g0 = 0; g0 + 16 <= n; g0 += 16) {
            int mem[16];
            for (int t = 0; t < 16; ++t) { mem[t] = order[g0 + t]; if (pos[mem[t]] >= 0) { remove(mem[t], pos[mem[t]]); pos[mem[t]] = -1; } }
            std::stable_sort(mem, mem + 16, [&](int a, int b) { return G.of_ptr[a + 1] - G.of_ptr[a] > G.of_ptr[b + 1] - G.of_ptr[b]; });
            bool used[16] = {false};
            for (int t = 0; t < 16; ++t) {
                const int x = mem[t];
                int cost[16] = {0};
                for (int q = G.of_ptr[x]; q < G.of_ptr[x + 1]; ++q) {
                    const int s2 = G.of_set[q];
                    const unsigned char *c = &cnt[(size_t)s2 * 16];
                    const int m2 = pairwise_only ? 255 : mx[s2];
                    for (int p = 0; p < 16; ++p) cost[p] += c[p] + (c[p] >= m2 ? 64 : 0);     // 64 > any sum of counts of one item's sets / position
                }
                int best = -1;
                for (int p = 0; p < 16; ++p) if (!used[p] && (best < 0 || cost[p] < cost[best])) best = p;
                used[best] = true; pos[x] = best; add(x, best);
            }
            for (int t = 0; t < 16; ++t) order[g0 + pos[mem[t]]] = mem[t];
        }
}
static void group_columns(const std::vector<int> &cl, int CH, std::vector<int> &order) {
    const int n = (int)cl.size();
    std::vector<int> sorted(n);
    for (int j = 0; j < n; ++j) sorted[j] = j;
    std::stable_sort(sorted.begin(), sorted.end(), [&](int a, int b) { return cl[a] > cl[b]; });
    std::vector<char> taken(n, 0);
    order.clear(); order.reserve(n);
    int first = 0;
    while ((int)order.size() < n) {
        while (first < n && taken[first]) ++first;
        const int Lmax = cl[sorted[first]];
        bool used[16] = {false};
        int got = 0;
        const size_t base = order.size();
        for (int q = first; q < n && got < 16 && cl[sorted[q]] >= Lmax - 1; ++q) {
            if (taken[q]) continue;
            const int j = sorted[q], c = (CH * (j & 3) + (j >> 2)) & 15;
            if (used[c]) continue;
            used[c] = true; taken[q] = 1; order.push_back(j); ++got;
        }
        for (int q = first; q < n && got < 16; ++q) if (!taken[q]) { taken[q] = 1; order.push_back(sorted[q]); ++got; }   // no distinct bank left
        std::stable_sort(order.begin() + base, order.end(), [&](int a, int b) { return cl[a] > cl[b]; });
    }
}
// Slot order of the rows (rord) and columns (cord) of one instance.  mode 0: descending stored length (stable); 1: bank-aware.
static void assign_slots(int ni, int mi, int nz, const int32_t *cp, const int32_t *ri, const std::vector<int> &rl, const std::vector<int> &cl,
                         int mode, int np_batch, std::vector<int> &rord, std::vector<int> &cord) {
    // one greedy sweep: further sweeps do not lower the wavefront count (1.74 x the conflict-free count after one sweep, 1.76 x after two,
    // tools/eval_placement.py; even 200 000 random improving swaps per instance only reach 1.65 x -- the rest is decided by which rows and
    // columns share a half-warp, not by the positions inside a group) and cost as much host time as the first
    static const int sweeps = getenv("LPBOX_PLACE_SWEEPS") ? atoi(getenv("LPBOX_PLACE_SWEEPS")) : 1;
    rord.resize(mi);
    for (int r = 0; r < mi; ++r) rord[r] = r;
    std::stable_sort(rord.begin(), rord.end(), [&](int a, int b2) { return rl[a] > rl[b2]; });
    if (mode == 0) {
        cord.resize(ni);
        for (int j = 0; j < ni; ++j) cord[j] = j;
        std::stable_sort(cord.begin(), cord.end(), [&](int a, int b2) { return cl[a] > cl[b2]; });
        return;
    }
    group_columns(cl, chain_stride(np_batch), cord);
    // row-compressed copy of the pattern (rows gather columns in E v)
    std::vector<int> rp(mi + 1, 0), cix(nz);
    for (int r = 0; r < mi; ++r) rp[r + 1] = rp[r] + rl[r];
    { std::vector<int> fill(rp.begin(), rp.end() - 1); for (int j = 0; j < ni; ++j) for (int k = cp[j]; k < cp[j + 1]; ++k) cix[fill[ri[k]]++] = j; }
    GatherSets gs;
    build_sets(rord, rp.data(), cix.data(), ni, gs);         // E v: half-warps of rows gather columns -> place the columns
    place_in_groups(cord, gs, sweeps);
    std::vector<int> cpi(cp, cp + ni + 1);
    build_sets(cord, cpi.data(), ri, mi, gs);                // E^T w: half-warps of columns gather rows -> place the rows
    place_in_groups(rord, gs, sweeps);
}
}  // namespace

// Diagnostic (host only, no device needed): shared-memory wavefronts of the operand gathers of one E v and one E^T w of an instance
// under the slot assignment `mode` (see assign_slots) -- a 64-bit access is served half a warp at a time, a half-warp takes as many
// wavefronts as the largest number of DISTINCT addresses that share one 8-byte bank pair.  out = {E v actual, E v ideal (one wavefront
// per half-warp and step), E^T w actual, E^T w ideal}.  cap = T * EPT of the kernel shape (512 for n <= 512).
extern "C" int lpbox_debug_gather_wavefronts(int m, int n, const int32_t *colptr, const int32_t *rowidx, int cap, int mode, int64_t *out) {
    if (m <= 0 || n <= 0 || !colptr || !rowidx || !out || cap < std::max(m, n)) return LPBOX_E_INVALID;
    if (colptr[0] != 0) return LPBOX_E_INVALID;
    for (int j = 0; j < n; ++j) if (colptr[j + 1] < colptr[j]) return LPBOX_E_INVALID;
    const int nz = colptr[n];
    std::vector<int> rl(m, 0), cl(n), rord, cord;
    for (int k = 0; k < nz; ++k) { if (rowidx[k] < 0 || rowidx[k] >= m) return LPBOX_E_INVALID; rl[rowidx[k]]++; }
    for (int j = 0; j < n; ++j) cl[j] = colptr[j + 1] - colptr[j];
    assign_slots(n, m, nz, colptr, rowidx, rl, cl, mode, std::max((n + 1) & ~1, (m + 1) & ~1), rord, cord);
    std::vector<int> rinv(m), cinv(n);
    for (int s = 0; s < m; ++s) rinv[rord[s]] = s;
    for (int s = 0; s < n; ++s) cinv[cord[s]] = s;
    std::vector<int> rp(m + 1, 0), cix(nz);
    for (int r = 0; r < m; ++r) rp[r + 1] = rp[r] + rl[r];
    { std::vector<int> fill(rp.begin(), rp.end() - 1); for (int j = 0; j < n; ++j) for (int k = colptr[j]; k < colptr[j + 1]; ++k) cix[fill[rowidx[k]]++] = j; }
    const int zero_word = zero_off(cap) / 8, t1_word = gather_base(cap) / 8;
    auto count = [&](const std::vector<int> &order, const int *ptr, const int *idx, const std::vector<int> &inv, int word0, int64_t &act, int64_t &ideal) {
        const int cnt = (int)order.size();
        act = ideal = 0;
        for (int s0 = 0; s0 < cnt; s0 += 32) {
            int W = 0;
            for (int s = s0; s < std::min(s0 + 32, cnt); ++s) W = std::max(W, ptr[order[s] + 1] - ptr[order[s]]);
            for (int k = 0; k < W; ++k)
                for (int h0 = s0; h0 < s0 + 32; h0 += 16) {
                    std::vector<int> addr[16];
                    for (int s = h0; s < h0 + 16; ++s) {
                        int w = zero_word;
                        if (s < cnt) { const int o = order[s]; if (ptr[o] + k < ptr[o + 1]) w = word0 + inv[idx[ptr[o] + k]]; }
                        auto &v = addr[w & 15];
                        if (std::find(v.begin(), v.end(), w) == v.end()) v.push_back(w);
                    }
                    size_t mx = 1;
                    for (auto &v : addr) mx = std::max(mx, v.size());
                    act += (int64_t)mx; ideal += 1;
                }
        }
    };
    std::vector<int> cpv(colptr, colptr + n + 1), riv(rowidx, rowidx + nz);
    count(rord, rp.data(), cix.data(), cinv, 0, out[0], out[1]);
    count(cord, cpv.data(), riv.data(), rinv, t1_word, out[2], out[3]);
    return 0;
}
namespace {
}  // namespace

static lpbox_batch *batch_create_impl(int device, int B, const int32_t *m, const int32_t *n, const int32_t *colptr_all,
                                      const int32_t *rowidx_all, const double *val_all, const double *b_all,
                                      const double *f_all, int hist_cap);
extern "C" lpbox_batch *lpbox_batch_create(int device, int B, const int32_t *m, const int32_t *n, const int32_t *colptr_all,
                                           const int32_t *rowidx_all, const double *val_all, const double *b_all,
                                           const double *f_all, int hist_cap) {
    try {   // no C++ exception may cross the C boundary
        return batch_create_impl(device, B, m, n, colptr_all, rowidx_all, val_all, b_all, f_all, hist_cap);
    } catch (const std::exception &e) {
        set_err(std::string("lpbox_batch_create: ") + e.what());
    } catch (...) {
        set_err("lpbox_batch_create: unknown exception");
    }
    return nullptr;
}
static lpbox_batch *batch_create_impl(int device, int B, const int32_t *m, const int32_t *n, const int32_t *colptr_all,
                                      const int32_t *rowidx_all, const double *val_all, const double *b_all,
                                      const double *f_all, int hist_cap) {
    if (B <= 0 || !m || !n || !colptr_all || !rowidx_all || !b_all || hist_cap < 0) { set_err("invalid argument"); return nullptr; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { set_err("no CUDA device (there is no CPU fallback)"); return nullptr; }
    if (device < 0 || device >= ndev) { set_err("bad device index"); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { set_err("cudaSetDevice failed"); return nullptr; }
    const bool dbg_t = getenv("LPBOX_DEBUG") != nullptr;                         // stage timings of the set-up on stderr
    auto t_prev = std::chrono::steady_clock::now();
    auto stage = [&](const char *what) {
        if (!dbg_t) return;
        const auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[lpbox] create: %-28s %7.1f ms\n", what, std::chrono::duration<double, std::milli>(t - t_prev).count());
        t_prev = t;
    };
    lpbox_batch *h = new lpbox_batch();
    h->device = device; h->B = B; h->hist_cap = hist_cap;
    h->n0.assign(n, n + B); h->m0.assign(m, m + B); h->nnz0.resize(B);
    h->off_n.assign(B + 1, 0); h->off_m.assign(B + 1, 0); h->off_pat.assign(B + 1, 0); h->off_val.assign(B + 1, 0);
    h->off_csr.assign(B + 1, 0); h->off_evr.assign(B + 1, 0); h->off_evc.assign(B + 1, 0);
    std::vector<int> rcap(B, 0), ccap(B, 0), maxcl(B, 0);
    std::vector<std::vector<uint16_t>> rperm_all(B), cperm_all(B);
    h->off_hist.assign(B + 1, 0); h->h_nnz_off.assign(B + 1, 0); h->h_cp_off.assign(B + 1, 0);
    // (1) cheap serial pass: sizes and the offsets that depend on n, m, nnz only
    std::vector<long long> b_in_off(B + 1, 0), f_in_off(B + 1, 0);
    for (int i = 0; i < B; ++i) {
        if (n[i] <= 0 || m[i] < 0) { set_err("instance with n <= 0"); delete h; return nullptr; }
        if (n[i] > 2048 || m[i] > 2048) { set_err("max(n, m) > 2048 is not supported by the on-chip kernel"); delete h; return nullptr; }
        const int32_t *cp = colptr_all + h->h_cp_off[i];
        // validate the column pointers BEFORE anything is sized from them
        bool ok_cp = cp[0] == 0;
        for (int j = 0; ok_cp && j < n[i]; ++j) ok_cp = cp[j + 1] >= cp[j];
        if (!ok_cp) { set_err("bad colptr (must start at 0 and be non-decreasing)"); delete h; return nullptr; }
        h->nnz0[i] = cp[n[i]];
        if (h->nnz0[i] > 65535) { set_err("nnz > 65535 is not supported by the on-chip kernel"); delete h; return nullptr; }
        h->h_cp_off[i + 1] = h->h_cp_off[i] + n[i] + 1;
        h->h_nnz_off[i + 1] = h->h_nnz_off[i] + h->nnz0[i];
        h->off_n[i + 1] = h->off_n[i] + ((n[i] + 1) & ~1);
        h->off_m[i + 1] = h->off_m[i] + ((m[i] + 1) & ~1);
        h->off_hist[i + 1] = h->off_hist[i] + (long long)hist_cap * n[i];
        b_in_off[i + 1] = b_in_off[i] + n[i]; f_in_off[i + 1] = f_in_off[i] + m[i];
        h->max_n = std::max(h->max_n, n[i]); h->max_m = std::max(h->max_m, m[i]); h->max_nnz = std::max(h->max_nnz, h->nnz0[i]);
    }
    stage("sizes / validation");
    // (2) per instance, in parallel on the host cores: SpMV work assignment -- slots in descending stored length, bank-aware
    //     (see above; LPBOX_PLAIN_SLOTS=1 keeps the plain stable sort) -- and the capacities of the sliced-ELL image
    std::atomic<int> err1(0);
    static const bool plain_slots = getenv("LPBOX_PLAIN_SLOTS") != nullptr;
    const int np_batch = std::max((h->max_n + 1) & ~1, (h->max_m + 1) & ~1);     // the window kernel's vector stride (configure())
    host_parallel_for(B, [&](int i) {
        const int ni = n[i], mi = m[i], nz = h->nnz0[i];
        const int32_t *cp = colptr_all + h->h_cp_off[i];
        const int32_t *ri = rowidx_all + h->h_nnz_off[i];
        std::vector<int> rl(mi, 0), cl(ni), rord, cord;
        for (int k = 0; k < nz; ++k) { if (ri[k] < 0 || ri[k] >= mi) { err1.store(1); return; } rl[ri[k]]++; }
        for (int j = 0; j < ni; ++j) cl[j] = cp[j + 1] - cp[j];
        assign_slots(ni, mi, nz, cp, ri, rl, cl, plain_slots ? 0 : 1, np_batch, rord, cord);
        rperm_all[i].resize(mi);
        for (int s0 = 0; s0 < mi; s0 += 32) { int w = 0; for (int s2 = s0; s2 < std::min(s0 + 32, mi); ++s2) w = std::max(w, rl[rord[s2]]); rcap[i] += w; }
        for (int r = 0; r < mi; ++r) rperm_all[i][r] = (uint16_t)rord[r];
        cperm_all[i].resize(ni);
        int mx = 0;
        for (int s0 = 0; s0 < ni; s0 += 32) { int w = 0; for (int s2 = s0; s2 < std::min(s0 + 32, ni); ++s2) w = std::max(w, cl[cord[s2]]); ccap[i] += w; mx = std::max(mx, w); }
        for (int j = 0; j < ni; ++j) cperm_all[i][j] = (uint16_t)cord[j];
        maxcl[i] = mx;
    });
    if (err1.load()) { set_err("row index out of range"); delete h; return nullptr; }
    stage("slot assignment (parallel)");
    // (3) serial: offsets that depend on the layouts
    for (int i = 0; i < B; ++i) {
        EllLayout EL = ell_layout(n[i], m[i], rcap[i], ccap[i]);
        CsrLayout PL = csr_layout(n[i], m[i], h->nnz0[i]);
        h->off_pat[i + 1] = h->off_pat[i] + EL.bytes;
        h->off_csr[i + 1] = h->off_csr[i] + PL.bytes;
        h->off_val[i + 1] = h->off_val[i] + ((h->nnz0[i] + 1) & ~1);
        h->off_evr[i + 1] = h->off_evr[i] + 32 * rcap[i];
        h->off_evc[i + 1] = h->off_evc[i] + 32 * ccap[i];
        h->max_csr = std::max(h->max_csr, PL.bytes);
        h->max_evr = std::max(h->max_evr, 32 * rcap[i]); h->max_evc = std::max(h->max_evc, 32 * ccap[i]);
        if (rcap[i] > 2047 || ccap[i] > 2047) { set_err("pattern too large for the on-chip kernel"); delete h; return nullptr; }
        h->max_pat = std::max(h->max_pat, EL.o_rperm);
        h->max_col_len = std::max(h->max_col_len, maxcl[i]);
        h->pat_bytes_i.push_back(EL.o_rperm); h->pat_head_i.push_back(EL.o_cidx);   // staged bytes with / without the column offsets
    }
    long long tot_nnz = h->h_nnz_off[B];
    // the host copies of the pattern (getters, feasibility checks) are made by a helper thread while the blobs are packed
    struct Joiner { std::thread t; ~Joiner() { if (t.joinable()) t.join(); } } copier;
    copier.t = std::thread([h, colptr_all, rowidx_all, tot_nnz, B]() {
        h->h_colptr.assign(colptr_all, colptr_all + h->h_cp_off[B]);
        h->h_rowidx.assign(rowidx_all, rowidx_all + tot_nnz);
    });
    std::vector<double> ones;
    if (val_all) {
        h->h_val.assign(val_all, val_all + tot_nnz);
        h->all_unit = true;
        for (long long k = 0; k < tot_nnz; ++k) if (val_all[k] != 1.0) { h->all_unit = false; break; }
    }
    if (h->all_unit && h->max_col_len > 255) {   // the unit kernel indexes its 1/diag table with 8-bit column lengths: use the general path
        h->all_unit = false;
        if (!val_all) { ones.assign((size_t)tot_nnz, 1.0); val_all = ones.data(); h->h_val = ones; }
    }
    stage("layout offsets / host copies");
    // build pattern blobs (+ values in both orders) on the host
    // The big staging buffers are allocated UNINITIALISED and every instance clears / fills its own part inside the parallel loop
    // (a value-initialised std::vector would zero ~340 MB for 10 000 instances on one core first: 70 ms of the set-up).
    struct RawBuf {
        unsigned char *p = nullptr; size_t n = 0;
        explicit RawBuf(size_t bytes) : p(static_cast<unsigned char *>(::operator new(std::max<size_t>(bytes, 1)))), n(bytes) {}
        ~RawBuf() { ::operator delete(p); }
        RawBuf(const RawBuf &) = delete; RawBuf &operator=(const RawBuf &) = delete;
        unsigned char *data() { return p; } size_t size() const { return n; }
    };
    RawBuf pat((size_t)h->off_pat[B]), csr((size_t)h->off_csr[B]);
    std::vector<double> val_r, val_c;
    if (!h->all_unit) { val_r.assign((size_t)h->off_val[B], 0.0); val_c.assign((size_t)h->off_val[B], 0.0); }
    RawBuf fraw(sizeof(double) * (size_t)h->off_m[B]), braw(sizeof(double) * (size_t)h->off_n[B]);
    double *const fvec = reinterpret_cast<double *>(fraw.data()), *const bvec = reinterpret_cast<double *>(braw.data());
    std::vector<InstState> st(B);
    std::atomic<int> err2(0);
    host_parallel_for(B, [&](int i) {
        const int ni = n[i], mi = m[i], nz = h->nnz0[i];
        const long long boff = b_in_off[i], foff = f_in_off[i];
        const int32_t *cp = colptr_all + h->h_cp_off[i];
        const int32_t *ri = rowidx_all + h->h_nnz_off[i];
        const double *va = val_all ? val_all + h->h_nnz_off[i] : nullptr;
        CsrLayout PL = csr_layout(ni, mi, nz);
        unsigned char *blob = csr.data() + h->off_csr[i];
        memset(blob, 0, (size_t)(h->off_csr[i + 1] - h->off_csr[i]));
        {
            EllLayout EL = ell_layout(ni, mi, rcap[i], ccap[i]);
            unsigned char *eb = pat.data() + h->off_pat[i];
            memset(eb, 0, (size_t)(h->off_pat[i + 1] - h->off_pat[i]));
            memcpy(eb + EL.o_rperm, rperm_all[i].data(), sizeof(uint16_t) * (size_t)mi);
            memcpy(eb + EL.o_cperm, cperm_all[i].data(), sizeof(uint16_t) * (size_t)ni);
        }
        uint16_t *rowptr = (uint16_t *)(blob + PL.o_rowptr), *colptr = (uint16_t *)(blob + PL.o_colptr);
        uint16_t *colidx = (uint16_t *)(blob + PL.o_colidx), *rowidx = (uint16_t *)(blob + PL.o_rowidx);
        std::vector<int> rcount(mi + 1, 0);
        for (int j = 0; j < ni; ++j) {
            if (cp[j] > cp[j + 1] || cp[0] != 0) { err2.store(1); return; }
            colptr[j] = (uint16_t)cp[j];
            for (int k = cp[j]; k < cp[j + 1]; ++k) {
                if (ri[k] < 0 || ri[k] >= mi || (k > cp[j] && ri[k] <= ri[k - 1])) {
                    err2.store(2); return;
                }
                rowidx[k] = (uint16_t)ri[k];
                rcount[ri[k] + 1]++;
            }
        }
        colptr[ni] = (uint16_t)nz;
        for (int r = 0; r < mi; ++r) rcount[r + 1] += rcount[r];
        for (int r = 0; r <= mi; ++r) rowptr[r] = (uint16_t)rcount[r];
        std::vector<int> pos(rcount.begin(), rcount.end() - 1);
        for (int j = 0; j < ni; ++j)
            for (int k = cp[j]; k < cp[j + 1]; ++k) {
                int q = pos[ri[k]]++;
                colidx[q] = (uint16_t)j;
                if (!h->all_unit) { val_r[(size_t)h->off_val[i] + q] = va[k]; val_c[(size_t)h->off_val[i] + k] = va[k]; }
            }
        memcpy(bvec + h->off_n[i], b_all + boff, sizeof(double) * (size_t)ni);
        for (long long k = h->off_n[i] + ni; k < h->off_n[i + 1]; ++k) bvec[k] = 0.0;                  // stride padding
        if (f_all) memcpy(fvec + h->off_m[i], f_all + foff, sizeof(double) * (size_t)mi);
        else for (int r = 0; r < mi; ++r) fvec[h->off_m[i] + r] = 1.0;                                // f = 1 (LP.cpp:2522)
        for (long long k = h->off_m[i] + mi; k < h->off_m[i + 1]; ++k) fvec[k] = 1.0;
        InstState &s = st[i];
        memset(&s, 0, sizeof(s));
        s.n0 = s.n = ni; s.m0 = s.m = mi; s.nnz0 = s.nnz = nz; s.rcap = rcap[i]; s.ccap = ccap[i]; s.unit = h->all_unit ? 1 : 0; s.std_obj = 1.0; s.rhoUpdated = 1;
    });
    copier.t.join();
    if (err2.load()) { set_err(err2.load() == 1 ? "bad colptr" : "row indices must be in range and strictly ascending within each column"); delete h; return nullptr; }
    h->h_st = st;
    lpbox_params lp; lpbox_params_lp(&lp);
    h->pr.stop_threshold = lp.stop_threshold; h->pr.std_threshold = lp.std_threshold; h->pr.max_iters = lp.max_iters;
    h->pr.initial_rho = lp.initial_rho; h->pr.rho_change_step = lp.rho_change_step; h->pr.gamma_val = lp.gamma_val;
    h->pr.learning_fact = lp.learning_fact; h->pr.history_size = (int)lp.history_size; h->pr.gamma_factor = lp.gamma_factor;
    h->pr.pcg_tol = lp.pcg_tol; h->pr.pcg_maxiters = lp.pcg_maxiters; h->pr.guard_first_iter = 1; h->pr.alpha_bailout = 1;

    bool ok = true;
    auto A = [&](cudaError_t e) { if (e != cudaSuccess) { if (ok) set_err(cudaGetErrorString(e)); ok = false; } };
    A(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    A(cudaEventCreate(&h->ev0)); A(cudaEventCreate(&h->ev1));
    size_t NN = (size_t)h->off_n[B], MM = (size_t)h->off_m[B];
    stage("pattern blobs (parallel)");
    A(h->d_off_n.alloc(B + 1)); A(h->d_off_m.alloc(B + 1)); A(h->d_off_pat.alloc(B + 1)); A(h->d_off_val.alloc(B + 1));
    A(h->d_off_csr.alloc(B + 1)); A(h->d_off_evr.alloc(B + 1)); A(h->d_off_evc.alloc(B + 1)); A(h->d_csr.alloc((size_t)h->off_csr[B]));
    A(h->d_off_hist.alloc(B + 1)); A(h->d_off_vec.alloc(B + 1));
    A(h->d_x.alloc(NN)); A(h->d_y1.alloc(NN)); A(h->d_y2.alloc(NN)); A(h->d_z1.alloc(NN)); A(h->d_z2.alloc(NN));
    A(h->d_b.alloc(NN)); A(h->d_Pd.alloc(NN)); A(h->d_Esq.alloc(NN)); A(h->d_ret_val.alloc(NN)); A(h->d_vec.alloc(NN));
    A(h->d_left.alloc(NN)); A(h->d_ret_idx.alloc(NN));
    A(h->d_y3.alloc(MM)); A(h->d_z4.alloc(MM)); A(h->d_f.alloc(MM));
    A(h->d_pat.alloc((size_t)h->off_pat[B]));
    if (!h->all_unit) {
        A(h->d_val_r.alloc((size_t)h->off_val[B])); A(h->d_val_c.alloc((size_t)h->off_val[B]));
        A(h->d_ev_r.alloc((size_t)h->off_evr[B])); A(h->d_ev_c.alloc((size_t)h->off_evc[B])); A(h->d_r4v.alloc((size_t)h->off_evc[B]));
        if (ok) { A(cudaMemset(h->d_ev_r.p, 0, sizeof(double) * std::max<size_t>(h->d_ev_r.n, 1))); A(cudaMemset(h->d_ev_c.p, 0, sizeof(double) * std::max<size_t>(h->d_ev_c.n, 1)));
                  A(cudaMemset(h->d_r4v.p, 0, sizeof(double) * std::max<size_t>(h->d_r4v.n, 1))); }
    }
    A(h->d_hist.alloc((size_t)h->off_hist[B]));
    A(h->d_st.alloc(B)); A(h->d_counter.alloc(1 + SM_RANK_SLOTS)); A(h->d_num.alloc(B));
    A(h->d_pow.alloc((size_t)h->max_n + 1));
    if (!ok) { lpbox_batch_destroy(h); return nullptr; }
    stage("device allocation");
    std::vector<double> powtab((size_t)h->max_n + 1);
    for (int k = 0; k <= h->max_n; ++k) powtab[k] = pow((double)k, 1.0 / 2);   // std::pow(n, 1.0/p), LP.cpp:427 (host libm)
    auto H2D = [&](void *d, const void *s, size_t bytes) { if (bytes) { A(cudaMemcpyAsync(d, s, bytes, cudaMemcpyHostToDevice, h->stream)); h->h2d_bytes += (int64_t)bytes; } };
    H2D(h->d_off_n.p, h->off_n.data(), sizeof(long long) * (B + 1));
    H2D(h->d_off_m.p, h->off_m.data(), sizeof(long long) * (B + 1));
    H2D(h->d_off_pat.p, h->off_pat.data(), sizeof(long long) * (B + 1));
    H2D(h->d_off_val.p, h->off_val.data(), sizeof(long long) * (B + 1));
    H2D(h->d_off_csr.p, h->off_csr.data(), sizeof(long long) * (B + 1));
    H2D(h->d_off_evr.p, h->off_evr.data(), sizeof(long long) * (B + 1));
    H2D(h->d_off_evc.p, h->off_evc.data(), sizeof(long long) * (B + 1));
    H2D(h->d_csr.p, csr.data(), csr.size());
    H2D(h->d_off_hist.p, h->off_hist.data(), sizeof(long long) * (B + 1));
    H2D(h->d_b.p, bvec, sizeof(double) * NN);
    H2D(h->d_f.p, fvec, sizeof(double) * MM);
    H2D(h->d_pat.p, pat.data(), pat.size());
    if (!h->all_unit) { H2D(h->d_val_r.p, val_r.data(), sizeof(double) * val_r.size()); H2D(h->d_val_c.p, val_c.data(), sizeof(double) * val_c.size()); }
    H2D(h->d_st.p, st.data(), sizeof(InstState) * (size_t)B);
    H2D(h->d_pow.p, powtab.data(), sizeof(double) * powtab.size());
    A(cudaStreamSynchronize(h->stream));
    if (!ok) { lpbox_batch_destroy(h); return nullptr; }
    BatchView &v = h->bv;
    v.B = B; v.hist_cap = hist_cap;
    v.off_n = h->d_off_n.p; v.off_m = h->d_off_m.p; v.off_pat = h->d_off_pat.p; v.off_val = h->d_off_val.p; v.off_hist = h->d_off_hist.p;
    v.off_csr = h->d_off_csr.p; v.off_evr = h->d_off_evr.p; v.off_evc = h->d_off_evc.p; v.csr = h->d_csr.p; v.ev_r = h->d_ev_r.p; v.ev_c = h->d_ev_c.p;
    v.x = h->d_x.p; v.y1 = h->d_y1.p; v.y2 = h->d_y2.p; v.z1 = h->d_z1.p; v.z2 = h->d_z2.p; v.b = h->d_b.p; v.Pd = h->d_Pd.p; v.Esq = h->d_Esq.p;
    v.y3 = h->d_y3.p; v.z4 = h->d_z4.p; v.f = h->d_f.p; v.pat = h->d_pat.p;
    v.val_r = h->d_val_r.p; v.val_c = h->d_val_c.p; v.r4v = h->d_r4v.p; v.st = h->d_st.p; v.hist = h->d_hist.p;
    v.left_idx = h->d_left.p; v.ret_idx = h->d_ret_idx.p; v.ret_val = h->d_ret_val.p; v.pow_tab = h->d_pow.p;
    stage("H2D");
    if (configure(h) != 0) { lpbox_batch_destroy(h); return nullptr; }
    stage("configure");
    lp_setup_kernel<<<B, 128, 0, h->stream>>>(h->bv, h->pr, 4, 0);     // build the sliced-ELL images on the device
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(h->stream) != cudaSuccess) { set_err("ELL build kernel failed"); lpbox_batch_destroy(h); return nullptr; }
    h->launches += 1;
    stage("image build kernel");
    return h;
}

extern "C" void lpbox_batch_destroy(lpbox_batch *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    h->d_off_n.free_(); h->d_off_m.free_(); h->d_off_pat.free_(); h->d_off_val.free_(); h->d_off_hist.free_(); h->d_off_vec.free_();
    h->d_off_csr.free_(); h->d_off_evr.free_(); h->d_off_evc.free_(); h->d_csr.free_(); h->d_ev_r.free_(); h->d_ev_c.free_();
    h->d_x.free_(); h->d_y1.free_(); h->d_y2.free_(); h->d_z1.free_(); h->d_z2.free_(); h->d_b.free_(); h->d_Pd.free_(); h->d_Esq.free_();
    h->d_y3.free_(); h->d_z4.free_(); h->d_f.free_(); h->d_val_r.free_(); h->d_val_c.free_(); h->d_r4v.free_(); h->d_hist.free_();
    h->d_ret_val.free_(); h->d_pow.free_(); h->d_vec.free_(); h->d_pat.free_(); h->d_st.free_(); h->d_left.free_(); h->d_ret_idx.free_();
    h->d_counter.free_(); h->d_num.free_(); h->d_work.free_(); h->d_park.free_(); h->d_err.free_(); h->d_pinp.free_(); h->d_pscore.free_(); h->d_meta.free_(); h->d_ring.free_();
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream && h->own_stream) cudaStreamDestroy(h->stream);
    h->d_active.free_(); h->d_row_off.free_();
    delete h;
}

// mode 0: parity (default, bit-identical to the reference); mode 1: fast (tree reductions over all warps + FMA in the vector
// updates; iterates agree to rounding level only -- never used for parity claims)
extern "C" int lpbox_batch_set_mode(lpbox_batch *h, int mode) {
    if (!h || (mode != 0 && mode != 1)) return LPBOX_E_INVALID;
    h->fast = mode == 1;
    return 0;
}

// Optional extension (off by default = the reference's deter_fix_2): fix-to-one proposals are only applied where the fixed
// part of the solution stays feasible (see lp_guard_kernel).
extern "C" int lpbox_batch_set_fix_guard(lpbox_batch *h, int on) {
    if (!h) return LPBOX_E_INVALID;
    h->guard = on != 0;
    return 0;
}

extern "C" int lpbox_batch_set_params(lpbox_batch *h, const lpbox_params *p, int variant) {
    if (!h || !p) return LPBOX_E_INVALID;
    if (p->history_size < 2 || p->history_size > 16 || p->history_size != floor(p->history_size)) { set_err("history_size must be an integer in [2,16]"); return LPBOX_E_INVALID; }
    if (p->projection_lp != 2) { set_err("projection_lp must be 2 (the reference always uses the 2-norm, LP.cpp:423-428)"); return LPBOX_E_INVALID; }
    if (p->rho_change_step <= 0) return LPBOX_E_INVALID;
    h->pr.stop_threshold = p->stop_threshold; h->pr.std_threshold = p->std_threshold; h->pr.max_iters = p->max_iters;
    h->pr.initial_rho = p->initial_rho; h->pr.rho_change_step = p->rho_change_step; h->pr.gamma_val = p->gamma_val;
    h->pr.learning_fact = p->learning_fact; h->pr.history_size = (int)p->history_size; h->pr.gamma_factor = p->gamma_factor;
    h->pr.pcg_tol = p->pcg_tol; h->pr.pcg_maxiters = p->pcg_maxiters;
    h->pr.guard_first_iter = (variant & 1) ? 1 : 0; h->pr.alpha_bailout = (variant & 2) ? 1 : 0;
    return 0;
}

static void time_begin(lpbox_batch *h) { cudaEventRecord(h->ev0, h->stream); }
static int time_end(lpbox_batch *h) {
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaEventSynchronize(h->ev1));
    float ms = 0; CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_ms = ms;
    return 0;
}

extern "C" int lpbox_batch_init(lpbox_batch *h, const double *x0_all) {
    if (!h) return LPBOX_E_INVALID;
    CK(cudaSetDevice(h->device));
    // reset current sizes / patterns are only valid for a never-compacted batch
    for (int i = 0; i < h->B; ++i) if (h->inited && h->h_st[i].n != h->h_st[i].n0) { set_err("re-init after early fixing is not supported; create a new batch"); return LPBOX_E_INVALID; }
    if (x0_all) {
        std::vector<double> xs((size_t)h->off_n[h->B], 0.0);
        long long o = 0;
        for (int i = 0; i < h->B; ++i) { memcpy(xs.data() + h->off_n[i], x0_all + o, sizeof(double) * (size_t)h->n0[i]); o += h->n0[i]; }
        CK(cudaMemcpyAsync(h->d_x.p, xs.data(), sizeof(double) * xs.size(), cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    time_begin(h);
    lp_setup_kernel<<<h->B, 128, 0, h->stream>>>(h->bv, h->pr, 3, x0_all ? 1 : 0);
    CK(cudaGetLastError());
    h->launches += 1;
    { int rc = time_end(h); if (rc) return rc; }
    { int rc = sync_states(h); if (rc) return rc; }
    h->inited = true;
    return 1;   // ADMM_lp_iters_init returns 1 (LP.cpp:762)
}

// the plain loop (lpbox_batch_iters / _solve) also writes every iterate into the history ring (hist_cap iterations per call),
// readable with lpbox_batch_get_x_iters: what print_fix_info == 2 dumps to xiter/*.csv (LP.cpp:903-909)
extern "C" int lpbox_batch_set_record_history(lpbox_batch *h, int on) {
    if (!h) return LPBOX_E_INVALID;
    if (on && h->hist_cap <= 0) { set_err("create the batch with hist_cap > 0 to record the history"); return LPBOX_E_INVALID; }
    h->record_plain = on != 0;
    return 0;
}

extern "C" int lpbox_batch_iters(lpbox_batch *h, int iter_start, int iter_end, int32_t *ret) {
    if (!h || !h->inited) { set_err("call lpbox_batch_init first"); return LPBOX_E_INVALID; }
    CK(cudaSetDevice(h->device));
    time_begin(h);
    if (iter_start == 0 && iter_end > 0) {   // update_expression(0) inside iteration 0 (LP.cpp:833)
        lp_setup_kernel<<<h->B, 128, 0, h->stream>>>(h->bv, h->pr, 2, 0);
        CK(cudaGetLastError());
        h->launches += 1;
    }
    int rc = run_window(h, iter_start, iter_end, 0, h->B > 1 ? 1 : 0);
    if (rc) return rc;
    rc = time_end(h); if (rc) return rc;
    rc = sync_states(h); if (rc) return rc;
    if (ret) for (int i = 0; i < h->B; ++i) ret[i] = h->h_st[i].last_ret;
    return h->h_st[0].last_ret;
}

extern "C" int lpbox_batch_iters_l2f(lpbox_batch *h, int iter_start, int iter_end, const double *vec_all, const int32_t *num,
                                     int32_t *ret) {
    if (!h || !h->inited) { set_err("call lpbox_batch_init first"); return LPBOX_E_INVALID; }
    CK(cudaSetDevice(h->device));
    const int skip_done = h->B > 1 ? 1 : 0;
    bool any = false;
    if (num) for (int i = 0; i < h->B; ++i) if (num[i] != 0) any = true;
    if (any && !vec_all) { set_err("vec_all is NULL but some num[i] != 0"); return LPBOX_E_INVALID; }
    time_begin(h);
    if (any) {
        std::vector<long long> off_vec(h->B + 1, 0);
        for (int i = 0; i < h->B; ++i) off_vec[i + 1] = off_vec[i] + h->h_st[i].n;
        // validate: num[i] must equal the number of entries in {0,1} (the reference prints an error and corrupts memory otherwise)
        for (int i = 0; i < h->B; ++i) {
            if (num[i] == 0) continue;
            int c = 0;
            const double *v = vec_all + off_vec[i];
            for (int k = 0; k < h->h_st[i].n; ++k) if (v[k] == 1.0 || v[k] == 0.0) c++;
            if (c != num[i]) { set_err("num[i] does not match the number of fixed entries in vec"); return LPBOX_E_INVALID; }
        }
        CK(cudaMemcpyAsync(h->d_off_vec.p, off_vec.data(), sizeof(long long) * (h->B + 1), cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->d_vec.p, vec_all, sizeof(double) * (size_t)off_vec[h->B], cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->d_num.p, num, sizeof(int) * (size_t)h->B, cudaMemcpyHostToDevice, h->stream));
    } else {
        CK(cudaMemsetAsync(h->d_num.p, 0, sizeof(int) * (size_t)h->B, h->stream));
    }
    int np = (h->max_n + 1) & ~1, mp = (h->max_m + 1) & ~1;
    int val_elems = h->all_unit ? 0 : ((h->max_nnz + 1) & ~1);
    lp_fix_kernel<<<h->B, FIX_T, h->fix_smem, h->stream>>>(h->bv, h->pr, h->d_vec.p, h->d_off_vec.p, h->d_num.p, skip_done, np, mp,
                                                          h->max_csr, val_elems);
    CK(cudaGetLastError());
    h->launches += 1;
    if (iter_start == 0 && iter_end > 0) {   // `if(iter==0) update_expression(0)` (LP.cpp:1380-1381)
        lp_setup_kernel<<<h->B, 128, 0, h->stream>>>(h->bv, h->pr, 2, 0);
        CK(cudaGetLastError());
        h->launches += 1;
    }
    int rc = run_window(h, iter_start, iter_end, 1, skip_done);
    if (rc) return rc;
    rc = time_end(h); if (rc) return rc;
    rc = sync_states(h); if (rc) return rc;
    if (ret) for (int i = 0; i < h->B; ++i) ret[i] = h->h_st[i].last_ret;
    return h->h_st[0].last_ret;
}

extern "C" int lpbox_batch_solve(lpbox_batch *h, int max_iters, lpbox_log_row *log) {
    if (!h || !h->inited) { set_err("call lpbox_batch_init first"); return LPBOX_E_INVALID; }
    CK(cudaSetDevice(h->device));
    time_begin(h);
    lp_setup_kernel<<<h->B, 128, 0, h->stream>>>(h->bv, h->pr, 2, 0);
    CK(cudaGetLastError());
    h->launches += 1;
    // Sliced (breadth-first) work queue: opt-in.  Measured on 4144 instances: 870 instances/s with slices of 250 or 500 iterations
    // against 927 with one instance per CTA until it stops -- when every instance advances at the same pace the launch ends with
    // ALL the long instances still alive and too few of them to fill the GPU, which costs more than the depth-first tail.
    const int slice_env = getenv("LPBOX_SLICE") ? atoi(getenv("LPBOX_SLICE")) : 0;
    int rc = run_window(h, 0, max_iters, 0, 1, h->record_plain ? 0 : slice_env);
    if (rc) return rc;
    rc = time_end(h); if (rc) return rc;
    rc = sync_states(h); if (rc) return rc;
    if (log) return lpbox_batch_results(h, log, nullptr, 0);
    return 0;
}

extern "C" int lpbox_batch_size(const lpbox_batch *h) { return h ? h->B : LPBOX_E_INVALID; }
#define CHK_I(h, i) if (!(h) || (i) < 0 || (i) >= (h)->B) return LPBOX_E_INVALID
extern "C" int lpbox_batch_get_n(lpbox_batch *h, int i) { CHK_I(h, i); return h->h_st[i].n; }
extern "C" int lpbox_batch_get_m(lpbox_batch *h, int i) { CHK_I(h, i); return h->h_st[i].m; }
extern "C" int lpbox_batch_get_org_n(lpbox_batch *h, int i) { CHK_I(h, i); return h->h_st[i].n0; }
extern "C" int lpbox_batch_get_iter(lpbox_batch *h, int i) { CHK_I(h, i); return h->h_st[i].iter; }
extern "C" double lpbox_batch_cal_obj(lpbox_batch *h, int i) {
    if (!h || i < 0 || i >= h->B) return NAN;
    const InstState &s = h->h_st[i];
    return s.n != 0 ? s.sum_fix_obj + s.cur_obj : s.sum_fix_obj;   // LP.cpp:1630-1642
}
extern "C" double lpbox_batch_get_cur_bin_obj(lpbox_batch *h, int i) { if (!h || i < 0 || i >= h->B) return NAN; return h->h_st[i].cur_obj; }

static int d2h(lpbox_batch *h, void *dst, const void *src, size_t bytes) {
    if (!bytes) return 0;
    h->d2h_bytes += (int64_t)bytes;
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int lpbox_batch_get_state(lpbox_batch *h, int i, double *x, double *y1, double *y2, double *z1, double *z2, double *y3, double *z4) {
    CHK_I(h, i);
    const InstState &s = h->h_st[i];
    size_t nb = sizeof(double) * (size_t)s.n, mb = sizeof(double) * (size_t)s.m;
    long long on = h->off_n[i], om = h->off_m[i];
    int rc = 0;
    if (x) rc |= d2h(h, x, h->d_x.p + on, nb);
    if (y1) rc |= d2h(h, y1, h->d_y1.p + on, nb);
    if (y2) rc |= d2h(h, y2, h->d_y2.p + on, nb);
    if (z1) rc |= d2h(h, z1, h->d_z1.p + on, nb);
    if (z2) rc |= d2h(h, z2, h->d_z2.p + on, nb);
    if (y3) rc |= d2h(h, y3, h->d_y3.p + om, mb);
    if (z4) rc |= d2h(h, z4, h->d_z4.p + om, mb);
    return rc ? LPBOX_E_CUDA : 0;
}

// left_idx (LP.h:246): original ids of the variables that are still free, n_cur entries (diagnostics of the early-fixing loop)
extern "C" int lpbox_batch_get_left_idx(lpbox_batch *h, int i, int32_t *out) {
    CHK_I(h, i);
    if (!out) return LPBOX_E_INVALID;
    return d2h(h, out, h->d_left.p + h->off_n[i], sizeof(int) * (size_t)h->h_st[i].n) ? LPBOX_E_CUDA : h->h_st[i].n;
}

extern "C" int lpbox_batch_get_final_x_sol(lpbox_batch *h, int i, double *out) {
    CHK_I(h, i);
    if (!out) return LPBOX_E_INVALID;
    return d2h(h, out, h->d_x.p + h->off_n[i], sizeof(double) * (size_t)h->h_st[i].n) ? LPBOX_E_CUDA : h->h_st[i].n;
}

// get_x_sol (LP.cpp:1648-1665)
static int assemble_x(lpbox_batch *h, int i, double *out) {
    const InstState &s = h->h_st[i];
    long long on = h->off_n[i];
    std::vector<int> ridx(s.n_ret), left(s.n);
    std::vector<double> rval(s.n_ret), x(s.n);
    if (d2h(h, ridx.data(), h->d_ret_idx.p + on, sizeof(int) * (size_t)s.n_ret)) return LPBOX_E_CUDA;
    if (d2h(h, rval.data(), h->d_ret_val.p + on, sizeof(double) * (size_t)s.n_ret)) return LPBOX_E_CUDA;
    if (d2h(h, left.data(), h->d_left.p + on, sizeof(int) * (size_t)s.n)) return LPBOX_E_CUDA;
    if (d2h(h, x.data(), h->d_x.p + on, sizeof(double) * (size_t)s.n)) return LPBOX_E_CUDA;
    for (int q = 0; q < s.n_ret; ++q) out[ridx[q]] = rval[q];
    for (int q = 0; q < s.n; ++q) out[left[q]] = (x[q] >= 0.5) ? 1.0 : 0.0;
    return s.n0;
}
extern "C" int lpbox_batch_get_x_sol(lpbox_batch *h, int i, double *out) {
    CHK_I(h, i);
    if (!out) return LPBOX_E_INVALID;
    return assemble_x(h, i, out);
}

extern "C" int lpbox_batch_get_x_iters(lpbox_batch *h, int i, int ws, double *out) {
    CHK_I(h, i);
    if (!out || ws <= 0) return LPBOX_E_INVALID;
    const InstState &s = h->h_st[i];
    int rows = s.xit_rows, cols = std::min(std::min(s.xit_cols, ws), h->hist_cap);
    std::vector<double> buf((size_t)std::max(cols, 0) * s.n0);
    if (cols > 0 && d2h(h, buf.data(), h->d_hist.p + h->off_hist[i], sizeof(double) * buf.size())) return LPBOX_E_CUDA;
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < ws; ++c) out[(size_t)r * ws + c] = (c < cols) ? buf[(size_t)c * s.n0 + r] : 0.0;   // x_iters is zero-initialised (LP.cpp:1113)
    return rows;
}

extern "C" int lpbox_batch_check_infeasible_lpbox(lpbox_batch *h, int i) {   // LP.cpp:1577-1591: current E, relaxed x
    CHK_I(h, i);
    const InstState &s = h->h_st[i];
    CsrLayout PL = csr_layout(s.n0, s.m0, s.nnz0);
    std::vector<unsigned char> blob(PL.bytes);
    std::vector<double> x(s.n), vr;
    if (d2h(h, blob.data(), h->d_csr.p + h->off_csr[i], blob.size())) return LPBOX_E_CUDA;
    if (d2h(h, x.data(), h->d_x.p + h->off_n[i], sizeof(double) * (size_t)s.n)) return LPBOX_E_CUDA;
    if (!s.unit) { vr.resize(s.nnz); if (d2h(h, vr.data(), h->d_val_r.p + h->off_val[i], sizeof(double) * (size_t)s.nnz)) return LPBOX_E_CUDA; }
    const uint16_t *rowptr = (const uint16_t *)(blob.data() + PL.o_rowptr), *colidx = (const uint16_t *)(blob.data() + PL.o_colidx);
    int inf = 0;
    for (int r = 0; r < s.m; ++r) {
        double acc = 0.0;
        for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) acc = acc + (s.unit ? 1.0 : vr[k]) * x[colidx[k]];
        if (!(acc <= 1.0)) inf++;
    }
    return inf;
}

extern "C" int lpbox_batch_check_infeasible_l2f(lpbox_batch *h, int i) {     // LP.cpp:1593-1612: original E, assembled binary x
    CHK_I(h, i);
    const InstState &s = h->h_st[i];
    std::vector<double> xs(s.n0, 0.0);
    int rc = assemble_x(h, i, xs.data());
    if (rc < 0) return rc;
    std::vector<double> acc(s.m0, 0.0);
    const int *cp = h->h_colptr.data() + h->h_cp_off[i];
    const int *ri = h->h_rowidx.data() + h->h_nnz_off[i];
    const double *va = h->h_val.empty() ? nullptr : h->h_val.data() + h->h_nnz_off[i];
    for (int j = 0; j < s.n0; ++j)
        for (int k = cp[j]; k < cp[j + 1]; ++k) acc[ri[k]] = acc[ri[k]] + (va ? va[k] : 1.0) * xs[j];
    int inf = 0;
    for (int r = 0; r < s.m0; ++r) if (!(acc[r] <= 1.0)) inf++;
    return inf;
}

extern "C" int lpbox_batch_results(lpbox_batch *h, lpbox_log_row *log, uint8_t *x_bits, int row_stride_bytes) {
    if (!h) return LPBOX_E_INVALID;
    int rc = sync_states(h); if (rc) return rc;
    size_t NN = (size_t)h->off_n[h->B];
    std::vector<double> x(NN), rval(NN);
    std::vector<int> left(NN), ridx(NN);
    if (d2h(h, x.data(), h->d_x.p, sizeof(double) * NN) || d2h(h, rval.data(), h->d_ret_val.p, sizeof(double) * NN) ||
        d2h(h, left.data(), h->d_left.p, sizeof(int) * NN) || d2h(h, ridx.data(), h->d_ret_idx.p, sizeof(int) * NN)) return LPBOX_E_CUDA;
    host_parallel_for(h->B, [&](int i) {
        std::vector<double> xs, acc;
        const InstState &s = h->h_st[i];
        long long on = h->off_n[i];
        xs.assign(s.n0, 0.0);
        for (int q = 0; q < s.n_ret; ++q) xs[ridx[on + q]] = rval[on + q];
        for (int q = 0; q < s.n; ++q) xs[left[on + q]] = (x[on + q] >= 0.5) ? 1.0 : 0.0;
        if (x_bits) {
            uint8_t *row = x_bits + (size_t)i * row_stride_bytes;
            memset(row, 0, row_stride_bytes);
            for (int j = 0; j < s.n0 && (j >> 3) < row_stride_bytes; ++j) if (xs[j] >= 0.5) row[j >> 3] |= (uint8_t)(1u << (j & 7));
        }
        if (log) {
            acc.assign(s.m0, 0.0);
            const int *cp = h->h_colptr.data() + h->h_cp_off[i];
            const int *ri = h->h_rowidx.data() + h->h_nnz_off[i];
            const double *va = h->h_val.empty() ? nullptr : h->h_val.data() + h->h_nnz_off[i];
            for (int j = 0; j < s.n0; ++j)
                for (int k = cp[j]; k < cp[j + 1]; ++k) acc[ri[k]] = acc[ri[k]] + (va ? va[k] : 1.0) * xs[j];
            int inf = 0;
            for (int r = 0; r < s.m0; ++r) if (!(acc[r] <= 1.0)) inf++;
            lpbox_log_row &L = log[i];
            L.iters = (int32_t)s.admm_iters; L.status = s.status; L.cg_iters = s.cg_iters;
            L.obj = s.n != 0 ? s.sum_fix_obj + s.cur_obj : s.sum_fix_obj; L.cur_bin_obj = s.cur_obj; L.n_left = s.n; L.infeasible = inf;
        }
    });
    return 0;
}

extern "C" double lpbox_batch_last_kernel_ms(const lpbox_batch *h) { return h ? h->last_ms : -1.0; }
// ---- device-resident window loop (no host copies of iterates / scores / fix vectors) ---------------------------------
extern "C" int lpbox_batch_set_stream(lpbox_batch *h, void *cuda_stream) {
    if (!h) return LPBOX_E_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)cuda_stream; h->own_stream = false;
    return 0;
}
extern "C" void *lpbox_batch_hist_dev(lpbox_batch *h) { return h ? (void *)h->d_hist.p : nullptr; }

extern "C" int64_t lpbox_batch_policy_input_dev(lpbox_batch *h, int ws, float *out_dev, int64_t capacity_rows) {
    if (!h || !h->inited || ws <= 0 || ws > h->hist_cap) { set_err("policy_input: bad arguments (ws must be <= hist_cap)"); return LPBOX_E_INVALID; }
    CK(cudaSetDevice(h->device));
    h->h_active.clear(); h->h_row_off.assign(1, 0);
    int max_rows = 0;
    for (int i = 0; i < h->B; ++i) {
        const InstState &s = h->h_st[i];
        if (s.done || s.n == 0) continue;
        h->h_active.push_back(i);
        h->h_row_off.push_back(h->h_row_off.back() + s.n);
        max_rows = std::max(max_rows, s.n);
    }
    const int na = (int)h->h_active.size();
    const int64_t rows = h->h_row_off.back();
    if (!out_dev || na == 0) return rows;
    if (rows > capacity_rows) { set_err("policy_input: output buffer too small"); return LPBOX_E_INVALID; }
    if (!h->d_active.p) { CK(h->d_active.alloc(h->B)); CK(h->d_row_off.alloc((size_t)h->B + 1)); }
    CK(cudaMemcpyAsync(h->d_active.p, h->h_active.data(), sizeof(int) * (size_t)na, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_row_off.p, h->h_row_off.data(), sizeof(long long) * (size_t)(na + 1), cudaMemcpyHostToDevice, h->stream));
    dim3 grid((max_rows + 31) / 32, na), block(32, 8);
    lp_policy_input_kernel<<<grid, block, 0, h->stream>>>(h->bv, h->d_active.p, h->d_row_off.p, ws, out_dev, nullptr);
    CK(cudaGetLastError());
    h->launches += 1;
    return rows;
}

extern "C" int lpbox_batch_apply_scores_dev(lpbox_batch *h, const float *scores_dev, double hi, double lo, int min_fix) {
    if (!h || !h->inited || !scores_dev) return LPBOX_E_INVALID;
    CK(cudaSetDevice(h->device));
    const int na = (int)h->h_active.size();
    CK(cudaMemsetAsync(h->d_num.p, 0, sizeof(int) * (size_t)h->B, h->stream));
    if (na > 0) {
        lp_threshold_kernel<<<na, 256, 0, h->stream>>>(h->bv, h->d_active.p, h->d_row_off.p, scores_dev, hi, lo, min_fix, h->d_vec.p,
                                                      h->d_off_vec.p, h->d_num.p, nullptr);
        CK(cudaGetLastError());
        h->launches += 1;
        if (h->guard) {
            lp_guard_kernel<<<na, 256, sizeof(unsigned long long) * (size_t)h->max_m, h->stream>>>(h->bv, h->d_active.p, h->d_row_off.p, scores_dev, min_fix,
                                                                                          h->d_vec.p, h->d_num.p, nullptr);
            CK(cudaGetLastError());
            h->launches += 1;
        }
    }
    h->dev_fix_pending = true;
    return na;
}

extern "C" int lpbox_batch_iters_l2f_dev(lpbox_batch *h, int iter_start, int iter_end) {
    if (!h || !h->inited) { set_err("call lpbox_batch_init first"); return LPBOX_E_INVALID; }
    CK(cudaSetDevice(h->device));
    time_begin(h);
    if (!h->dev_fix_pending) CK(cudaMemsetAsync(h->d_num.p, 0, sizeof(int) * (size_t)h->B, h->stream));
    h->dev_fix_pending = false;
    int np = (h->max_n + 1) & ~1, mp = (h->max_m + 1) & ~1;
    int val_elems = h->all_unit ? 0 : ((h->max_nnz + 1) & ~1);
    lp_fix_kernel<<<h->B, FIX_T, h->fix_smem, h->stream>>>(h->bv, h->pr, h->d_vec.p, h->d_off_vec.p, h->d_num.p, 1, np, mp, h->max_csr, val_elems);
    CK(cudaGetLastError());
    h->launches += 1;
    if (iter_start == 0 && iter_end > 0) {
        lp_setup_kernel<<<h->B, 128, 0, h->stream>>>(h->bv, h->pr, 2, 0);
        CK(cudaGetLastError());
        h->launches += 1;
    }
    int rc = run_window(h, iter_start, iter_end, 1, 1);
    if (rc) return rc;
    rc = time_end(h); if (rc) return rc;
    rc = sync_states(h); if (rc) return rc;
    int active = 0;
    for (int i = 0; i < h->B; ++i) if (!h->h_st[i].done && h->h_st[i].n != 0) active++;
    return active;
}


// ---- the whole early-fixing loop behind one call (LP.trainer:510-545) ------------------------------------------------------
// for w in range(max_iter / ws): ADMM_lp_iters_l2f(ws*w, ws*(w+1), fix vector of the previous window); stop when every instance
// returned 1; policy on the window's iterate history; deter_fix_2 (p > hi -> 1, p < lo -> 0, <= min_fix fixes -> none).
// Everything is enqueued on the handle's stream; the active list and the row offsets of the policy input are built on the
// device, and the host reads back 24 bytes per window (how many rows the policy has to score).
extern "C" int lpbox_batch_solve_l2f(lpbox_batch *h, lpbox_policy *policy, int ws, int max_iter, double hi, double lo, int min_fix,
                                     lpbox_log_row *log, uint8_t *x_bits, int row_stride_bytes, lpbox_l2f_stats *stats) {
    if (!h || !h->inited || !policy || ws <= 0 || max_iter < ws) { set_err("solve_l2f: bad arguments / call lpbox_batch_init first"); return LPBOX_E_INVALID; }
    if (ws > h->hist_cap) { set_err("solve_l2f: create the batch with hist_cap >= ws"); return LPBOX_E_INVALID; }
    if (ws % h->pr.rho_change_step != 0) { set_err("solve_l2f: ws must be a multiple of rho_change_step (LP.cpp:1392-1405)"); return LPBOX_E_INVALID; }
    CK(cudaSetDevice(h->device));
    const long long cap_rows = h->off_n[h->B];
    if (!h->d_pinp.p || h->pinp_ws != ws) {
        h->d_pinp.free_(); h->d_pscore.free_();
        CK(h->d_pinp.alloc((size_t)cap_rows * (size_t)ws)); CK(h->d_pscore.alloc((size_t)cap_rows));
        h->pinp_ws = ws;
    }
    if (!h->d_meta.p) CK(h->d_meta.alloc(1));
    if (!h->d_active.p) { CK(h->d_active.alloc(h->B)); CK(h->d_row_off.alloc((size_t)h->B + 1)); }
    const int np = (h->max_n + 1) & ~1, mp = (h->max_m + 1) & ~1;
    const int val_elems = h->all_unit ? 0 : ((h->max_nnz + 1) & ~1);
    lpbox_l2f_stats st{};
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, h->stream));
    CK(cudaMemsetAsync(h->d_num.p, 0, sizeof(int) * (size_t)h->B, h->stream));
    int rc = 0;
    const bool dbg_t = getenv("LPBOX_DEBUG") != nullptr;          // phase timings (fix / window / scan+gather / policy+threshold) on stderr
    std::vector<cudaEvent_t> evs;
    auto mark = [&]() { if (dbg_t) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, h->stream); evs.push_back(e); } };
    for (int w = 0; w < max_iter / ws && rc == 0; ++w) {
        mark();
        lp_fix_kernel<<<h->B, FIX_T, h->fix_smem, h->stream>>>(h->bv, h->pr, h->d_vec.p, h->d_off_vec.p, h->d_num.p, 1, np, mp, h->max_csr, val_elems);
        if (w == 0) lp_setup_kernel<<<h->B, 128, 0, h->stream>>>(h->bv, h->pr, 2, 0);     // `if(iter==0) update_expression(0)` (LP.cpp:1380-1381)
        h->launches += (w == 0) ? 2 : 1;
        mark();
        rc = run_window(h, ws * w, ws * (w + 1), 1, 1);
        if (rc) break;
        mark();
        st.windows++;
        lp_active_scan_kernel<<<1, 1024, 0, h->stream>>>(h->bv, h->d_active.p, h->d_row_off.p, h->d_meta.p);
        h->launches += 1;
        L2fMeta meta{};
        CK(cudaMemcpyAsync(&meta, h->d_meta.p, sizeof(meta), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        h->d2h_bytes += (int64_t)sizeof(meta);
        if (meta.n_active == 0 || meta.rows == 0) break;
        dim3 grid((unsigned)((meta.max_n + 31) / 32), (unsigned)meta.n_active), block(32, 8);
        lp_policy_input_kernel<<<grid, block, 0, h->stream>>>(h->bv, h->d_active.p, h->d_row_off.p, ws, h->d_pinp.p, h->d_meta.p);
        h->launches += 1;
        mark();
        rc = lpbox_policy_forward_dev(policy, (void *)h->stream, h->d_pinp.p, meta.rows, h->d_pscore.p);
        if (rc) break;
        st.policy_rows += meta.rows;
        CK(cudaMemsetAsync(h->d_num.p, 0, sizeof(int) * (size_t)h->B, h->stream));
        lp_threshold_kernel<<<(unsigned)meta.n_active, 256, 0, h->stream>>>(h->bv, h->d_active.p, h->d_row_off.p, h->d_pscore.p, hi, lo, min_fix,
                                                                         h->d_vec.p, h->d_off_vec.p, h->d_num.p, h->d_meta.p);
        h->launches += 1;
        if (h->guard) {
            lp_guard_kernel<<<(unsigned)meta.n_active, 256, sizeof(unsigned long long) * (size_t)h->max_m, h->stream>>>(
                h->bv, h->d_active.p, h->d_row_off.p, h->d_pscore.p, min_fix, h->d_vec.p, h->d_num.p, h->d_meta.p);
            h->launches += 1;
        }
    }
    if (rc == 0 && cudaGetLastError() != cudaSuccess) { set_err("solve_l2f: kernel launch failed"); rc = LPBOX_E_CUDA; }
    cudaEventRecord(e1, h->stream);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    if (dbg_t && rc == 0) {
        // marks per complete window: [fix | window | scan + gather | policy + threshold (+ guard)]; the last window may stop after its scan
        double ph[4] = {0, 0, 0, 0};
        for (size_t i = 0; i + 1 < evs.size(); ++i) { float t = 0; cudaEventElapsedTime(&t, evs[i], evs[i + 1]); ph[i & 3] += t; }
        if (!evs.empty()) { float t = 0; cudaEventElapsedTime(&t, evs.back(), e1); ph[(evs.size() - 1) & 3] += t; }
        fprintf(stderr, "[lpbox] solve_l2f: %d windows, %.1f ms: fix %.1f, window kernel %.1f, scan + input gather %.1f, policy + thresholds %.1f\n",
                (int)st.windows, ms, ph[0], ph[1], ph[2], ph[3]);
    }
    for (cudaEvent_t e : evs) cudaEventDestroy(e);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (rc) return rc;
    h->last_ms = ms; st.device_ms = ms;
    rc = sync_states(h); if (rc) return rc;
    if (stats) *stats = st;
    if (log || x_bits) return lpbox_batch_results(h, log, x_bits, row_stride_bytes);
    return 0;
}

extern "C" int lpbox_batch_config(const lpbox_batch *h, int32_t *out4) {
    if (!h || !out4) return LPBOX_E_INVALID;
    out4[0] = h->grid; out4[1] = (int32_t)h->smem; out4[2] = h->threads; out4[3] = (int32_t)h->fix_smem;
    return 0;
}
extern "C" int64_t lpbox_batch_h2d_bytes(const lpbox_batch *h) { return h ? h->h2d_bytes : -1; }
extern "C" int64_t lpbox_batch_d2h_bytes(const lpbox_batch *h) { return h ? h->d2h_bytes : -1; }
extern "C" int64_t lpbox_batch_launch_count(const lpbox_batch *h) { return h ? h->launches : -1; }

// ---- file format of the reference (readFile / readSparseMat / readDenseVec, LP.cpp:2407-2545) ----------------------
extern "C" void lpbox_free(void *p) { free(p); }
extern "C" int lpbox_read_instance(const char *root, int i, int k, int j, int32_t *m_out, int32_t *n_out, int32_t **colptr_out,
                                   int32_t **rowidx_out, double **val_out, double **b_out) {
    if (!root || !m_out || !n_out || !colptr_out || !rowidx_out || !val_out || !b_out) return LPBOX_E_INVALID;
    char pc[1024], pb[1024];
    snprintf(pc, sizeof(pc), "%s/instance/%d_%d/instance_%d_C.txt", root, k, j, i);   // :2462
    snprintf(pb, sizeof(pb), "%s/instance/%d_%d/instance_%d_b.txt", root, k, j, i);   // :2463
    FILE *fc = fopen(pc, "r");
    if (!fc) { set_err(std::string("cannot open ") + pc); return LPBOX_E_IO; }
    struct Trip { int r, c; double v; };
    std::vector<Trip> tr;
    int row, col, max_row = 0, max_col = 0;
    double val;
    while (fscanf(fc, "%d,%d,%lf\n", &row, &col, &val) == 3) {                       // :2422
        max_row = std::max(max_row, row); max_col = std::max(max_col, col);
        tr.push_back({row - 1, col - 1, (k == 2) ? -1.0 * val : val});                // :2433-2436
    }
    fclose(fc);
    const int m = max_row, n = max_col;
    // setFromTriplets: column-major, duplicates summed, inner indices ascending
    std::stable_sort(tr.begin(), tr.end(), [](const Trip &a, const Trip &b) { return a.c != b.c ? a.c < b.c : a.r < b.r; });
    std::vector<Trip> u;
    for (const Trip &t : tr) {
        if (t.r < 0 || t.c < 0) { set_err("bad triplet"); return LPBOX_E_IO; }
        if (!u.empty() && u.back().r == t.r && u.back().c == t.c) u.back().v += t.v; else u.push_back(t);
    }
    int32_t *cp = (int32_t *)calloc((size_t)n + 1, sizeof(int32_t));
    int32_t *ri = (int32_t *)malloc(sizeof(int32_t) * std::max<size_t>(u.size(), 1));
    double *va = (double *)malloc(sizeof(double) * std::max<size_t>(u.size(), 1));
    double *b = (double *)malloc(sizeof(double) * std::max<size_t>((size_t)n, 1));
    for (size_t q = 0; q < u.size(); ++q) { cp[u[q].c + 1]++; ri[q] = u[q].r; va[q] = u[q].v; }
    for (int c = 0; c < n; ++c) cp[c + 1] += cp[c];
    FILE *fb = fopen(pb, "r");
    if (!fb) { free(cp); free(ri); free(va); free(b); set_err(std::string("cannot open ") + pb); return LPBOX_E_IO; }
    for (int q = 0; q < n; ++q) {
        if (fscanf(fb, "%lf\n", &b[q]) != 1) { fclose(fb); free(cp); free(ri); free(va); free(b); set_err("error when reading dense vector"); return LPBOX_E_IO; }
        b[q] = -1.0 * b[q];                                                          // :2520
    }
    fclose(fb);
    *m_out = m; *n_out = n; *colptr_out = cp; *rowidx_out = ri; *val_out = va; *b_out = b;
    return 0;
}
