// Host side + C ABI of the batched unconstrained (segmentation) solver.  See include/lpbox_b200.h.
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/lpbox_b200.h"
#include "seg_kernels.cuh"

using namespace lpb;

void lpbox_set_error(const std::string &s);   // lp_batch.cu

#define SCK(call)                                                                      \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) { lpbox_set_error(std::string(#call) + ": " + cudaGetErrorString(e_)); return LPBOX_E_CUDA; } \
    } while (0)

template <typename Tp>
struct SBuf {
    Tp *p = nullptr;
    cudaError_t alloc(size_t count) { return cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(Tp)); }
    void free_() { if (p) cudaFree(p); p = nullptr; }
};

struct lpbox_seg_batch {
    int device = 0, B = 0, hist_cap = 0;
    bool compact = false;        // A stored as (int16 column distance, int8 value): see SegFmt in seg_kernels.cuh
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<int> n0, nnz0;
    std::vector<long long> off_n, off_nnz, off_hist;
    std::vector<double> cconst;
    Params pr{};
    SegView sv{};
    SBuf<long long> d_off_n, d_off_nnz, d_off_hist;
    SBuf<double> vecs[11], d_b[2], d_val[2], d_hist, d_ret_val, d_powv, d_powtab, d_vec, d_b_org, d_val_org;
    SBuf<int> d_rp[2], d_ci[2], d_left, d_ret_idx, d_counter, d_kidx, d_cnt, d_num, d_rp_org, d_ci_org;
    SBuf<long long> d_off_vec;
    SBuf<uint4> d_ell_c;                        // compact format: row image the ADMM kernel reads (seg_ell_build_kernel)
    SBuf<uint2> d_ell_a;
    SBuf<int> d_ell_flag;
    int max_n = 0;
    SBuf<SegInst> d_st;
    std::vector<SegInst> h_st;
    bool inited = false;
    int grid = 0;
    int threads = 0;                            // launch shape of seg_admm_kernel (SegCfg): 256, 192 or 160
    size_t smem_admm = 0;
    size_t smem = 0;
    double last_ms = 0;
    int64_t launches = 0, h2d_bytes = 0, d2h_bytes = 0;
};

// ---- graph builder (SEG.cpp:55-81, :144-248, :727-758) ------------------------------------------------------------------
// Grey image (row-major uint8, nr x nc) -> A = D - W (row-compressed, <= 7 stored entries per row, explicit zeros kept),
// b = U2 - U1, c = sum U1.  Restated from the reference: I = grey/263; unary costs from a Gaussian background model
// (mean 0.6) and a two-Gaussian foreground model (means 0.2, 0.2), sigma 0.1, rounded half away from zero; pairwise
// weights round(3 exp(-(Ip - Iq)^2 / sigma_img)) over the offsets (a,b) in {-1,0,1}^2 with a != b (a 6-neighbourhood),
// sigma_img = sample standard deviation of the image; intensities are looked up through the column-major flattening while
// p, q are row-major indices (the reference's index mismatch, reproduced literally).
static double eigen_sum(const double *v, long n) {   // Eigen's linear-vectorised redux order (sum)
    long a2 = (n / 4) * 4, a1 = (n / 2) * 2;
    if (!a1) return n ? v[0] : 0.0;
    double p00 = v[0], p01 = v[1];
    if (a1 > 2) {
        double p10 = v[2], p11 = v[3];
        for (long i = 4; i < a2; i += 4) { p00 += v[i]; p01 += v[i + 1]; p10 += v[i + 2]; p11 += v[i + 3]; }
        p00 += p10; p01 += p11;
        if (a1 > a2) { p00 += v[a2]; p01 += v[a2 + 1]; }
    }
    double res = p00 + p01;
    for (long i = a1; i < n; ++i) res += v[i];
    return res;
}

extern "C" int lpbox_seg_build_graph(const uint8_t *pixels, int nr, int nc, int32_t *rowptr, int32_t *colidx, double *val, double *b,
                                     double *c_out) {
    if (!pixels || nr <= 0 || nc <= 0 || !rowptr || !colidx || !val || !b || !c_out) return LPBOX_E_INVALID;
    const int n = nr * nc;
    std::vector<double> I(n), v(n), tmp(n);
    for (int k = 0; k < n; ++k) I[k] = (double)pixels[k] / 263.0;                                  // SEG.cpp:727
    for (int c = 0; c < nc; ++c) for (int r = 0; r < nr; ++r) v[(size_t)c * nr + r] = I[(size_t)r * nc + c];   // vectorize :46-53
    const double sigma = 0.1, bg = 0.6, f1 = 0.2, f2 = 0.2;                                        // :734-737
    const double cst = log(2.0 * 3.14159265358979323846) / 2.0 + log(sigma);
    for (int k = 0; k < n; ++k) {
        const double ab = pow(v[k] - bg, 2.0) / (2 * sigma * sigma) + cst;                         // :58
        const double aa = exp(-pow(v[k] - f1, 2.0) / (2 * sigma * sigma)) + exp(-pow(v[k] - f2, 2) / (2 * sigma * sigma));
        const double af = -log(aa + DBL_EPSILON) + cst + log(2.0);                                 // :61
        const double U1 = round(ab), U2 = round(af);                                               // :743
        b[k] = U2 - U1;
        tmp[k] = U1;
    }
    *c_out = eigen_sum(tmp.data(), n);
    const double mean = eigen_sum(v.data(), n) / (double)n;                                        // :181
    for (int k = 0; k < n; ++k) tmp[k] = (v[k] - mean) * (v[k] - mean);
    const double sig = sqrt(eigen_sum(tmp.data(), n) / (double)(n - 1));
    const int oa[7] = {-1, -1, 0, 0, 0, 1, 1}, ob[7] = {0, 1, -1, 0, 1, -1, 0};                    // ascending column order
    int q = 0;
    for (int i = 0; i < nr; ++i)
        for (int j = 0; j < nc; ++j) {
            const int p = i * nc + j;
            rowptr[p] = q;
            int dq = -1;
            for (int t = 0; t < 7; ++t) {
                const int a = oa[t], bo = ob[t];
                if (a == 0 && bo == 0) { dq = q; colidx[q] = p; val[q] = 0.0; q++; continue; }     // explicit diagonal :213-219
                if (i + a < 0 || i + a >= nr || j + bo < 0 || j + bo >= nc) continue;
                const int p2 = (i + a) * nc + (j + bo);
                const double i1 = I[(size_t)(p % nr) * nc + (p / nr)], i2 = I[(size_t)(p2 % nr) * nc + (p2 / nr)];   // :192-193
                const double wgt = round(3 * exp(-(pow(i1 - i2, 2.0) / sig)));                     // :196-205
                colidx[q] = p2; val[q] = -wgt; q++;
            }
            double wsum = 0.0;
            for (int e = rowptr[p]; e < q; ++e) wsum = wsum + (-val[e]) * 1.0;                     // We = -A 1  (:234-236)
            val[dq] = val[dq] + wsum;
        }
    rowptr[n] = q;
    return q;
}

static int seg_sync_states(lpbox_seg_batch *h) {
    h->d2h_bytes += (int64_t)(sizeof(SegInst) * (size_t)h->B);
    SCK(cudaMemcpyAsync(h->h_st.data(), h->d_st.p, sizeof(SegInst) * (size_t)h->B, cudaMemcpyDeviceToHost, h->stream));
    SCK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" void lpbox_seg_destroy(lpbox_seg_batch *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    h->d_off_n.free_(); h->d_off_nnz.free_(); h->d_off_hist.free_();
    for (auto &v : h->vecs) v.free_();
    for (int k = 0; k < 2; ++k) { h->d_b[k].free_(); h->d_val[k].free_(); h->d_rp[k].free_(); h->d_ci[k].free_(); }
    h->d_hist.free_(); h->d_ret_val.free_(); h->d_powv.free_(); h->d_powtab.free_(); h->d_vec.free_(); h->d_b_org.free_(); h->d_val_org.free_();
    h->d_kidx.free_(); h->d_cnt.free_(); h->d_num.free_(); h->d_rp_org.free_(); h->d_ci_org.free_(); h->d_off_vec.free_(); h->d_left.free_(); h->d_ret_idx.free_(); h->d_counter.free_(); h->d_st.free_();
    h->d_ell_c.free_(); h->d_ell_a.free_(); h->d_ell_flag.free_();
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

// allocation + kernel configuration shared by both constructors; CSR / b are filled afterwards (H2D or the device graph builder)
// Opt-in shared memory + resident CTAs per SM of one launch shape of seg_admm_kernel.
template <int T>
static cudaError_t seg_shape_query(bool compact, int *occ, size_t *smem) {
    *smem = sizeof(double) * (SegCfg<T>::BUF + 8 + 16);
    const int sm = (int)*smem;
    cudaError_t e;
    if (compact) {
        e = cudaFuncSetAttribute(seg_admm_kernel<true, T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(seg_admm_kernel<true, T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        int o2 = 0;
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, seg_admm_kernel<true, T, true>, T, *smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o2, seg_admm_kernel<true, T, false>, T, *smem);
        if (e == cudaSuccess) *occ = std::min(*occ, o2);
    } else {
        e = cudaFuncSetAttribute(seg_admm_kernel<false, T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, seg_admm_kernel<false, T, false>, T, *smem);
    }
    return e;
}

static lpbox_seg_batch *seg_new(int device, int B, const int32_t *n, const int *nnz, const double *c, int hist_cap, bool compact) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { lpbox_set_error("no CUDA device (there is no CPU fallback)"); return nullptr; }
    if (device < 0 || device >= ndev || cudaSetDevice(device) != cudaSuccess) { lpbox_set_error("bad device"); return nullptr; }
    lpbox_seg_batch *h = new lpbox_seg_batch();
    h->device = device; h->B = B; h->hist_cap = hist_cap; h->compact = compact;
    h->n0.assign(n, n + B); h->nnz0.assign(nnz, nnz + B); h->cconst.assign(B, 0.0);
    h->off_n.assign(B + 1, 0); h->off_nnz.assign(B + 1, 0); h->off_hist.assign(B + 1, 0);
    std::vector<double> powv(B);
    std::vector<SegInst> st(B);
    for (int i = 0; i < B; ++i) {
        h->off_n[i + 1] = h->off_n[i] + ((n[i] + 3) & ~3);
        h->off_nnz[i + 1] = h->off_nnz[i] + ((h->nnz0[i] + 3) & ~3);
        h->off_hist[i + 1] = h->off_hist[i] + (long long)hist_cap * n[i];
        if (c) h->cconst[i] = c[i];
        h->max_n = std::max(h->max_n, n[i]);
        powv[i] = pow((double)n[i], 1.0 / 2);
        SegInst &s = st[i];
        memset(&s, 0, sizeof(s));
        s.n0 = s.n = n[i]; s.nnz0 = s.nnz = h->nnz0[i]; s.std_obj = 1.0; s.rhoUpdated = 1; s.cconst = h->cconst[i];
    }
    const size_t NN = (size_t)h->off_n[B], ZZ = (size_t)h->off_nnz[B];
    h->h_st = st;
    lpbox_params sp; lpbox_params_seg(&sp);
    h->pr.stop_threshold = sp.stop_threshold; h->pr.std_threshold = sp.std_threshold; h->pr.max_iters = sp.max_iters;
    h->pr.initial_rho = sp.initial_rho; h->pr.rho_change_step = sp.rho_change_step; h->pr.gamma_val = sp.gamma_val;
    h->pr.learning_fact = sp.learning_fact; h->pr.history_size = (int)sp.history_size; h->pr.gamma_factor = sp.gamma_factor;
    h->pr.pcg_tol = sp.pcg_tol; h->pr.pcg_maxiters = sp.pcg_maxiters; h->pr.guard_first_iter = 0; h->pr.alpha_bailout = 0;
    bool ok = true;
    auto A = [&](cudaError_t e) { if (e != cudaSuccess) { if (ok) lpbox_set_error(cudaGetErrorString(e)); ok = false; } };
    A(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    A(cudaEventCreate(&h->ev0)); A(cudaEventCreate(&h->ev1));
    A(h->d_off_n.alloc(B + 1)); A(h->d_off_nnz.alloc(B + 1)); A(h->d_off_hist.alloc(B + 1));
    for (auto &v : h->vecs) A(v.alloc(NN));
    // element counts of the typed buffers that hold colidx / val: 2 + 1 bytes per entry in the compact format, 4 + 8 otherwise
    const size_t ZCI = compact ? (ZZ + 1) / 2 : ZZ, ZVA = compact ? (ZZ + 7) / 8 : ZZ;
    for (int k = 0; k < 2; ++k) { A(h->d_b[k].alloc(NN)); A(h->d_val[k].alloc(ZVA)); A(h->d_rp[k].alloc(NN + 4 * (size_t)B)); A(h->d_ci[k].alloc(ZCI)); }
    A(h->d_hist.alloc((size_t)h->off_hist[B])); A(h->d_ret_val.alloc(NN)); A(h->d_powv.alloc(B)); A(h->d_left.alloc(NN)); A(h->d_ret_idx.alloc(NN));
    A(h->d_counter.alloc(1)); A(h->d_st.alloc(B));
    A(h->d_kidx.alloc(NN)); A(h->d_cnt.alloc(NN)); A(h->d_num.alloc(B)); A(h->d_off_vec.alloc(B + 1)); A(h->d_vec.alloc(NN)); A(h->d_powtab.alloc((size_t)h->max_n + 1));
    if (compact) { A(h->d_ell_c.alloc(NN)); A(h->d_ell_a.alloc(NN)); A(h->d_ell_flag.alloc(1)); }
    A(h->d_rp_org.alloc(NN + 4 * (size_t)B)); A(h->d_ci_org.alloc(ZCI)); A(h->d_val_org.alloc(ZVA)); A(h->d_b_org.alloc(NN));
    if (!ok) { lpbox_seg_destroy(h); return nullptr; }
    auto H2D = [&](void *d, const void *s, size_t bytes) { if (bytes) { A(cudaMemcpyAsync(d, s, bytes, cudaMemcpyHostToDevice, h->stream)); h->h2d_bytes += (int64_t)bytes; } };
    H2D(h->d_off_n.p, h->off_n.data(), sizeof(long long) * (B + 1));
    H2D(h->d_off_nnz.p, h->off_nnz.data(), sizeof(long long) * (B + 1));
    H2D(h->d_off_hist.p, h->off_hist.data(), sizeof(long long) * (B + 1));
    H2D(h->d_powv.p, powv.data(), sizeof(double) * (size_t)B);
    std::vector<double> powtab((size_t)h->max_n + 1);
    for (int k = 0; k <= h->max_n; ++k) powtab[k] = pow((double)k, 1.0 / 2);     // std::pow(n, 1.0/p) from the host libm
    H2D(h->d_powtab.p, powtab.data(), sizeof(double) * powtab.size());
    H2D(h->d_st.p, st.data(), sizeof(SegInst) * (size_t)B);
    A(cudaStreamSynchronize(h->stream));
    if (!ok) { lpbox_seg_destroy(h); return nullptr; }
    SegView &v = h->sv;
    v.B = B; v.hist_cap = hist_cap; v.off_n = h->d_off_n.p; v.off_nnz = h->d_off_nnz.p; v.off_hist = h->d_off_hist.p;
    double **vp[11] = {&v.x, &v.y1, &v.y2, &v.z1, &v.z2, &v.md, &v.invd, &v.r, &v.p, &v.t, &v.w};
    for (int k = 0; k < 11; ++k) *vp[k] = h->vecs[k].p;
    for (int k = 0; k < 2; ++k) { v.b[k] = h->d_b[k].p; v.rowptr[k] = h->d_rp[k].p; v.colidx[k] = h->d_ci[k].p; v.val[k] = h->d_val[k].p; }
    v.ell_c = h->d_ell_c.p; v.ell_a = h->d_ell_a.p; v.use_ell = 0;
    v.kidx = h->d_kidx.p; v.cnt = h->d_cnt.p; v.pow_tab = h->d_powtab.p;
    v.st = h->d_st.p; v.hist = h->d_hist.p; v.left_idx = h->d_left.p; v.ret_idx = h->d_ret_idx.p; v.ret_val = h->d_ret_val.p; v.powv = h->d_powv.p;
    h->smem = sizeof(double) * (SEG_BUF_DOUBLES + 8 + 16);                     // set-up kernel
    int sms = 0;
    if (cudaFuncSetAttribute(seg_setup_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem) != cudaSuccess ||
        cudaFuncSetAttribute(seg_setup_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
        lpbox_set_error("seg kernel configuration failed"); lpbox_seg_destroy(h); return nullptr;
    }
    // Launch shape (SegCfg): an image is solved by one CTA from start to end, so a launch takes `waves` x (time of one image),
    // and one image advances in proportion to its CTA's staging threads.  Pick the shape with the smallest waves / staging threads.
    const int shapes[3] = {256, 192, 160};
    double best = 0.0;
    const char *force = getenv("LPBOX_SEG_T");                                  // experiments
    for (int k = 0; k < 3; ++k) {
        int occ = 0; size_t sm = 0;
        const cudaError_t e = shapes[k] == 256 ? seg_shape_query<256>(compact, &occ, &sm)
                            : shapes[k] == 192 ? seg_shape_query<192>(compact, &occ, &sm) : seg_shape_query<160>(compact, &occ, &sm);
        if (e != cudaSuccess || occ < 1) { lpbox_set_error("seg kernel configuration failed"); lpbox_seg_destroy(h); return nullptr; }
        const int slots = sms * occ;
        const double cost = (double)((B + slots - 1) / slots) / (shapes[k] - 32);
        if (force ? atoi(force) == shapes[k] : (h->threads == 0 || cost < best)) {
            best = cost; h->threads = shapes[k]; h->smem_admm = sm; h->grid = std::max(1, std::min(B, slots));
        }
    }
    if (h->threads == 0) { lpbox_set_error("LPBOX_SEG_T must be 256, 192 or 160"); lpbox_seg_destroy(h); return nullptr; }
    return h;
}

extern "C" lpbox_seg_batch *lpbox_seg_create_csr(int device, int B, const int32_t *n, const int32_t *rowptr_all, const int32_t *colidx_all,
                                                const double *val_all, const double *b_all, const double *c, int hist_cap) {
    if (B <= 0 || !n || !rowptr_all || !colidx_all || !val_all || !b_all || hist_cap < 0) { lpbox_set_error("invalid argument"); return nullptr; }
    std::vector<int> nnz(B);
    std::vector<long long> rp_off(B + 1, 0);
    long long zo = 0;
    for (int i = 0; i < B; ++i) {
        if (n[i] <= 0) { lpbox_set_error("n <= 0"); return nullptr; }
        const int32_t *rp = rowptr_all + rp_off[i];
        nnz[i] = rp[n[i]];
        rp_off[i + 1] = rp_off[i] + n[i] + 1;
        for (int r = 0; r < n[i]; ++r) {
            bool diag = false;
            if (rp[r] > rp[r + 1]) { lpbox_set_error("bad rowptr"); return nullptr; }
            for (int k = rp[r]; k < rp[r + 1]; ++k) {
                const int cc = colidx_all[zo + k];
                if (cc < 0 || cc >= n[i] || (k > rp[r] && cc <= colidx_all[zo + k - 1])) { lpbox_set_error("column indices must be in range and strictly ascending within each row"); return nullptr; }
                if (cc == r) diag = true;
            }
            if (!diag) { lpbox_set_error("every row of A must store its diagonal entry (explicit zero allowed), as the reference's graph builder does"); return nullptr; }
        }
        zo += nnz[i];
    }
    // compact storage when every value is an integer in int8 range and every stored column is < 2^15 rows from its row
    bool compact = true;
    {
        long long z = 0;
        for (int i = 0; i < B && compact; ++i) {
            const int32_t *rp = rowptr_all + rp_off[i];
            for (int r = 0; r < n[i] && compact; ++r)
                for (int k = rp[r]; k < rp[r + 1]; ++k) {
                    const double a = val_all[z + k];
                    const long long d = (long long)colidx_all[z + k] - r;
                    if (!(a >= -128.0 && a <= 127.0) || a != (double)(signed char)a || d < -32768 || d > 32767) { compact = false; break; }
                }
            z += nnz[i];
        }
    }
    lpbox_seg_batch *h = seg_new(device, B, n, nnz.data(), c, hist_cap, compact);
    if (!h) return nullptr;
    const size_t NN = (size_t)h->off_n[B], ZZ = (size_t)h->off_nnz[B];
    std::vector<int> rp_pack(NN + 4 * (size_t)B, 0), ci_pack(compact ? 0 : ZZ, 0);
    std::vector<double> va_pack(compact ? 0 : ZZ, 0.0), b_pack(NN, 0.0);
    std::vector<short> ci16(compact ? ZZ : 0, 0);
    std::vector<signed char> va8(compact ? ZZ : 0, 0);
    long long bo = 0;
    zo = 0;
    for (int i = 0; i < B; ++i) {
        const int ni = n[i], nz = nnz[i];
        memcpy(rp_pack.data() + h->off_n[i] + 4 * (size_t)i, rowptr_all + rp_off[i], sizeof(int) * ((size_t)ni + 1));
        if (compact) {
            const int32_t *rp = rowptr_all + rp_off[i];
            for (int r = 0; r < ni; ++r)
                for (int k = rp[r]; k < rp[r + 1]; ++k) {
                    ci16[h->off_nnz[i] + k] = (short)(colidx_all[zo + k] - r);
                    va8[h->off_nnz[i] + k] = (signed char)val_all[zo + k];
                }
        } else {
            memcpy(ci_pack.data() + h->off_nnz[i], colidx_all + zo, sizeof(int) * (size_t)nz);
            memcpy(va_pack.data() + h->off_nnz[i], val_all + zo, sizeof(double) * (size_t)nz);
        }
        memcpy(b_pack.data() + h->off_n[i], b_all + bo, sizeof(double) * (size_t)ni);
        zo += nz; bo += ni;
    }
    bool ok = true;
    auto H2D = [&](void *d, const void *s, size_t bytes) {
        if (bytes) { if (cudaMemcpyAsync(d, s, bytes, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) ok = false; h->h2d_bytes += (int64_t)bytes; }
    };
    H2D(h->d_rp[0].p, rp_pack.data(), sizeof(int) * rp_pack.size());
    const void *ci_src = compact ? (const void *)ci16.data() : (const void *)ci_pack.data();
    const void *va_src = compact ? (const void *)va8.data() : (const void *)va_pack.data();
    const size_t ci_bytes = (compact ? sizeof(short) : sizeof(int)) * ZZ, va_bytes = (compact ? sizeof(signed char) : sizeof(double)) * ZZ;
    H2D(h->d_ci[0].p, ci_src, ci_bytes);
    H2D(h->d_val[0].p, va_src, va_bytes);
    H2D(h->d_b[0].p, b_pack.data(), sizeof(double) * NN);
    H2D(h->d_rp_org.p, rp_pack.data(), sizeof(int) * rp_pack.size());
    H2D(h->d_ci_org.p, ci_src, ci_bytes);
    H2D(h->d_val_org.p, va_src, va_bytes);
    H2D(h->d_b_org.p, b_pack.data(), sizeof(double) * NN);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) ok = false;
    if (!ok) { lpbox_set_error("upload of the CSR arrays failed"); lpbox_seg_destroy(h); return nullptr; }
    return h;
}

// ---- device graph builder: the same construction as lpbox_seg_build_graph, one CTA per image ------------------
// All outputs are integer-valued (rounded costs / weights and their sums), so they do not depend on summation order; the
// image mean / standard deviation, which feed exp(), are nevertheless summed in the host builder's (Eigen's) order by 4 lanes.
struct SegBuildArgs {
    const uint8_t *pixels; const long long *pix_off; const int *nr, *nc;
    const long long *off_n, *off_nnz;
    int *rowptr; void *colidx; void *val; double *b, *scr_v, *scr_t;      // colidx / val in the format of SegFmt<CMP>
    SegInst *st;
    int *bad;          // set when a weight cannot be stored in the compact format (NaN weights of a constant image)
    double cst, den, log2v, bg, f1, f2;
};

__device__ double seg_build_eigen_sum4(const double *v, long n, double *s_part /*[4]*/) {   // called by all threads; result valid in thread 0
    const long a2 = (n / 4) * 4, a1 = (n / 2) * 2;
    if (threadIdx.x < 4 && a2 >= 4) {
        double acc = v[threadIdx.x];
        for (long i = 4 + threadIdx.x; i < a2; i += 4) acc = acc + v[i];
        s_part[threadIdx.x] = acc;
    }
    __syncthreads();
    double res = 0.0;
    if (threadIdx.x == 0) {
        if (!a1) res = n ? v[0] : 0.0;
        else {
            double p00, p01;
            if (a2 >= 4) { p00 = s_part[0] + s_part[2]; p01 = s_part[1] + s_part[3]; if (a1 > a2) { p00 += v[a2]; p01 += v[a2 + 1]; } }
            else { p00 = v[0]; p01 = v[1]; }
            res = p00 + p01;
            for (long i = a1; i < n; ++i) res += v[i];
        }
    }
    __syncthreads();
    return res;
}

template <bool CMP>
__global__ void __launch_bounds__(256) seg_build_graph_kernel(SegBuildArgs a) {
    using CI = typename SegFmt<CMP>::CI;
    using AV = typename SegFmt<CMP>::AV;
    __shared__ double s_part[4], s_bc[2], s_red[8];
    const int img = blockIdx.x, tid = threadIdx.x;
    const int nr = a.nr[img], nc = a.nc[img], n = nr * nc;
    const uint8_t *pix = a.pixels + a.pix_off[img];
    double *v = a.scr_v + a.off_n[img], *t = a.scr_t + a.off_n[img], *b = a.b + a.off_n[img];
    int *rp = a.rowptr + a.off_n[img] + 4 * (long long)img;
    CI *ci = reinterpret_cast<CI *>(a.colidx) + a.off_nnz[img];
    AV *va = reinterpret_cast<AV *>(a.val) + a.off_nnz[img];
    // unary costs in the column-major flattening (SEG.cpp:46-61, :727-743)
    double c_part = 0.0;
    for (int k = tid; k < n; k += 256) {
        const int c = k / nr, r = k - c * nr;
        const double x = (double)pix[(size_t)r * nc + c] / 263.0;
        v[k] = x;
        const double d0 = x - a.bg, d1 = x - a.f1, d2 = x - a.f2;
        const double ab = (d0 * d0) / a.den + a.cst;
        const double aa = exp(-(d1 * d1) / a.den) + exp(-(d2 * d2) / a.den);
        const double af = -log(aa + 2.220446049250313e-16) + a.cst + a.log2v;
        const double U1 = round(ab), U2 = round(af);
        b[k] = U2 - U1;
        c_part += U1;                                   // integers: exact in any order
    }
    for (int o = 16; o > 0; o >>= 1) c_part += __shfl_xor_sync(0xffffffffu, c_part, o);
    if ((tid & 31) == 0) s_red[tid >> 5] = c_part;
    __syncthreads();
    if (tid == 0) { double cs = 0.0; for (int w = 0; w < 8; ++w) cs += s_red[w]; a.st[img].cconst = cs; }
    __syncthreads();
    // image statistics (SEG.cpp:181-186): mean and sample standard deviation
    const double sum = seg_build_eigen_sum4(v, n, s_part);
    if (tid == 0) s_bc[0] = sum / (double)n;
    __syncthreads();
    const double mean = s_bc[0];
    for (int k = tid; k < n; k += 256) { const double d = v[k] - mean; t[k] = d * d; }
    __syncthreads();
    const double ss = seg_build_eigen_sum4(t, n, s_part);
    if (tid == 0) s_bc[1] = sqrt(ss / (double)(n - 1));
    __syncthreads();
    const double sig = s_bc[1];
    // pairwise weights, row-major p; offsets in ascending column order; explicit diagonal (SEG.cpp:144-248)
    const int oa[7] = {-1, -1, 0, 0, 0, 1, 1}, ob[7] = {0, 1, -1, 0, 1, -1, 0};
    for (int p = tid; p < n; p += 256) {
        const int i = p / nc, j = p - i * nc;
        const int up = i > 0, dn = i < nr - 1;
        // entries stored before row i, and before column j inside row i (closed form of the stencil's validity pattern)
        const long long rows_before = (long long)i * (3 * nc - 2) + (long long)(2 * nc - 1) * ((i > 0 ? i - 1 : 0) + (i < nr - 1 ? i : nr - 1));
        const int jm = j < nc - 1 ? j : nc - 1, jl = j > 0 ? j - 1 : 0;
        const int in_row = j + up * j + up * jm + jl + jm + dn * jl + dn * j;
        int q = (int)(rows_before + in_row);
        rp[p] = q;
        if (p == n - 1) rp[n] = q + 1 + up + (up && j < nc - 1) + (j > 0) + (j < nc - 1) + (dn && j > 0) + dn;
        const double i1 = (double)pix[(size_t)(p % nr) * nc + (p / nr)] / 263.0;           // the reference's index mismatch (:192-193)
        int dq = -1;
        double wsum = 0.0;
#pragma unroll
        for (int e = 0; e < 7; ++e) {
            const int ao = oa[e], bo = ob[e];
            if (ao == 0 && bo == 0) { dq = q; ci[q] = (CI)(CMP ? 0 : p); q++; continue; }
            if (i + ao < 0 || i + ao >= nr || j + bo < 0 || j + bo >= nc) continue;
            const int p2 = (i + ao) * nc + (j + bo);
            const double i2 = (double)pix[(size_t)(p2 % nr) * nc + (p2 / nr)] / 263.0;
            const double d = i1 - i2;
            const double wgt = round(3 * exp(-((d * d) / sig)));
            if (CMP && !(wgt >= 0.0 && wgt <= 3.0)) *a.bad = 1;            // NaN (constant image): not representable as int8
            ci[q] = (CI)(CMP ? p2 - p : p2); va[q] = (AV)(-wgt); q++;      // weights are 0..3: exact in int8
            wsum += wgt;                                // integers: exact in any order
        }
        va[dq] = (AV)wsum;                              // <= 18
    }
}

static lpbox_seg_batch *seg_create_images_fmt(int device, int B, const uint8_t *pixels_all, const int32_t *nr, const int32_t *nc,
                                              int hist_cap, bool compact, bool *unrepresentable) {
    if (B <= 0 || !pixels_all || !nr || !nc || hist_cap < 0) { lpbox_set_error("invalid argument"); return nullptr; }
    std::vector<long long> po(B + 1, 0);
    std::vector<int> ns(B), nnz(B);
    for (int i = 0; i < B; ++i) {
        if (nr[i] <= 0 || nc[i] <= 0 || (long long)nr[i] * nc[i] > (1 << 28)) { lpbox_set_error("bad image size"); return nullptr; }
        const long long r = nr[i], c = nc[i];
        ns[i] = (int)(r * c);
        po[i + 1] = po[i] + r * c;
        // stored entries: diagonal + the valid ones of the 6 offsets (-1,0) (-1,1) (0,-1) (0,1) (1,-1) (1,0)
        nnz[i] = (int)(r * c + 2 * (r - 1) * c + 2 * r * (c - 1) + 2 * (r - 1) * (c - 1));
    }
    lpbox_seg_batch *h = seg_new(device, B, ns.data(), nnz.data(), nullptr, hist_cap, compact);
    if (!h) return nullptr;
    const size_t NN = (size_t)h->off_n[B], ZZ = (size_t)h->off_nnz[B];
    SBuf<uint8_t> d_pix; SBuf<long long> d_po; SBuf<int> d_nr, d_nc, d_bad;
    const size_t ci_sz = compact ? sizeof(short) : sizeof(int), va_sz = compact ? sizeof(signed char) : sizeof(double);
    int bad = 0;
    bool ok = d_bad.alloc(1) == cudaSuccess && d_pix.alloc((size_t)po[B]) == cudaSuccess && d_po.alloc(B + 1) == cudaSuccess && d_nr.alloc(B) == cudaSuccess && d_nc.alloc(B) == cudaSuccess;
    auto C_ = [&](cudaError_t e) { if (e != cudaSuccess) { if (ok) lpbox_set_error(std::string("seg graph builder: ") + cudaGetErrorString(e)); ok = false; } };
    if (ok) {
        C_(cudaMemcpyAsync(d_pix.p, pixels_all, (size_t)po[B], cudaMemcpyHostToDevice, h->stream));
        C_(cudaMemcpyAsync(d_po.p, po.data(), sizeof(long long) * (B + 1), cudaMemcpyHostToDevice, h->stream));
        C_(cudaMemcpyAsync(d_nr.p, nr, sizeof(int) * (size_t)B, cudaMemcpyHostToDevice, h->stream));
        C_(cudaMemcpyAsync(d_nc.p, nc, sizeof(int) * (size_t)B, cudaMemcpyHostToDevice, h->stream));
        h->h2d_bytes += (int64_t)po[B] + (int64_t)sizeof(long long) * (B + 1) + 2 * (int64_t)sizeof(int) * B;
        C_(cudaMemsetAsync(h->d_rp[0].p, 0, sizeof(int) * (NN + 4 * (size_t)B), h->stream));
        C_(cudaMemsetAsync(h->d_ci[0].p, 0, ci_sz * ZZ, h->stream));
        C_(cudaMemsetAsync(h->d_val[0].p, 0, va_sz * ZZ, h->stream));
        C_(cudaMemsetAsync(d_bad.p, 0, sizeof(int), h->stream));
        C_(cudaMemsetAsync(h->d_b[0].p, 0, sizeof(double) * NN, h->stream));
        SegBuildArgs a;
        a.pixels = d_pix.p; a.pix_off = d_po.p; a.nr = d_nr.p; a.nc = d_nc.p; a.off_n = h->d_off_n.p; a.off_nnz = h->d_off_nnz.p;
        a.rowptr = h->d_rp[0].p; a.colidx = h->d_ci[0].p; a.val = h->d_val[0].p; a.b = h->d_b[0].p; a.bad = d_bad.p;
        a.scr_v = h->vecs[7].p; a.scr_t = h->vecs[8].p; a.st = h->d_st.p;
        const double sigma = 0.1;                                                               // SEG.cpp:734-737
        a.cst = log(2.0 * 3.14159265358979323846) / 2.0 + log(sigma); a.den = 2 * sigma * sigma; a.log2v = log(2.0);
        a.bg = 0.6; a.f1 = 0.2; a.f2 = 0.2;
        if (compact) seg_build_graph_kernel<true><<<B, 256, 0, h->stream>>>(a);
        else seg_build_graph_kernel<false><<<B, 256, 0, h->stream>>>(a);
        C_(cudaGetLastError());
        h->launches++;
        C_(cudaMemcpyAsync(h->d_rp_org.p, h->d_rp[0].p, sizeof(int) * (NN + 4 * (size_t)B), cudaMemcpyDeviceToDevice, h->stream));
        C_(cudaMemcpyAsync(h->d_ci_org.p, h->d_ci[0].p, ci_sz * ZZ, cudaMemcpyDeviceToDevice, h->stream));
        C_(cudaMemcpyAsync(h->d_val_org.p, h->d_val[0].p, va_sz * ZZ, cudaMemcpyDeviceToDevice, h->stream));
        C_(cudaMemcpyAsync(&bad, d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        C_(cudaMemcpyAsync(h->d_b_org.p, h->d_b[0].p, sizeof(double) * NN, cudaMemcpyDeviceToDevice, h->stream));
        C_(cudaMemcpyAsync(h->h_st.data(), h->d_st.p, sizeof(SegInst) * (size_t)B, cudaMemcpyDeviceToHost, h->stream));
        C_(cudaStreamSynchronize(h->stream));
        for (int i = 0; i < B && ok; ++i) h->cconst[i] = h->h_st[i].cconst;
    } else {
        lpbox_set_error("seg graph builder: out of device memory");
    }
    d_pix.free_(); d_po.free_(); d_nr.free_(); d_nc.free_(); d_bad.free_();
    if (!ok) { lpbox_seg_destroy(h); return nullptr; }
    if (unrepresentable) *unrepresentable = bad != 0;
    if (bad) { lpbox_seg_destroy(h); return nullptr; }
    return h;
}

extern "C" lpbox_seg_batch *lpbox_seg_create_images(int device, int B, const uint8_t *pixels_all, const int32_t *nr, const int32_t *nc,
                                                   int hist_cap) {
    if (B <= 0 || !pixels_all || !nr || !nc || hist_cap < 0) { lpbox_set_error("invalid argument"); return nullptr; }
    bool compact = true, bad = false;
    for (int i = 0; i < B; ++i) if (nc[i] + 1 > 32767) compact = false;       // the int16 column distance would overflow
    if (compact) {
        lpbox_seg_batch *h = seg_create_images_fmt(device, B, pixels_all, nr, nc, hist_cap, true, &bad);
        if (h || !bad) return h;                                             // bad: a constant image (NaN weights) -> general format
    }
    return seg_create_images_fmt(device, B, pixels_all, nr, nc, hist_cap, false, nullptr);
}

// the graph the device builder produced for image i (tests / inspection): arrays sized n+1, nnz, nnz, n
// colidx / val of problem i as int32 / fp64, from the buffers `ci_dev` / `va_dev` (either storage format)
static int seg_fetch_matrix(lpbox_seg_batch *h, int i, const void *ci_dev, const void *va_dev, const int32_t *rowptr, int32_t *colidx, double *val) {
    const int n = h->n0[i], nz = h->nnz0[i];
    const size_t oz = (size_t)h->off_nnz[i];
    if (!h->compact) {
        SCK(cudaMemcpy(colidx, (const int *)ci_dev + oz, sizeof(int) * (size_t)nz, cudaMemcpyDeviceToHost));
        SCK(cudaMemcpy(val, (const double *)va_dev + oz, sizeof(double) * (size_t)nz, cudaMemcpyDeviceToHost));
        return 0;
    }
    std::vector<short> c16((size_t)nz);
    std::vector<signed char> v8((size_t)nz);
    SCK(cudaMemcpy(c16.data(), (const short *)ci_dev + oz, sizeof(short) * (size_t)nz, cudaMemcpyDeviceToHost));
    SCK(cudaMemcpy(v8.data(), (const signed char *)va_dev + oz, (size_t)nz, cudaMemcpyDeviceToHost));
    for (int r = 0; r < n; ++r)
        for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) { colidx[k] = r + c16[k]; val[k] = (double)v8[k]; }
    return 0;
}

extern "C" int lpbox_seg_get_graph(lpbox_seg_batch *h, int i, int32_t *rowptr, int32_t *colidx, double *val, double *b, double *c) {
    if (!h || i < 0 || i >= h->B) return LPBOX_E_INVALID;
    if (cudaSetDevice(h->device) != cudaSuccess) return LPBOX_E_CUDA;
    const int n = h->n0[i], nz = h->nnz0[i];
    SCK(cudaMemcpy(rowptr, h->d_rp_org.p + h->off_n[i] + 4 * (size_t)i, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToHost));
    if (int rc = seg_fetch_matrix(h, i, h->d_ci_org.p, h->d_val_org.p, rowptr, colidx, val)) return rc;
    SCK(cudaMemcpy(b, h->d_b_org.p + h->off_n[i], sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    if (c) *c = h->cconst[i];
    return nz;
}

extern "C" int lpbox_seg_set_params(lpbox_seg_batch *h, const lpbox_params *p) {
    if (!h || !p) return LPBOX_E_INVALID;
    if (p->history_size < 2 || p->history_size > 16 || p->rho_change_step <= 0) return LPBOX_E_INVALID;
    h->pr.stop_threshold = p->stop_threshold; h->pr.std_threshold = p->std_threshold; h->pr.max_iters = p->max_iters;
    h->pr.initial_rho = p->initial_rho; h->pr.rho_change_step = p->rho_change_step; h->pr.gamma_val = p->gamma_val;
    h->pr.learning_fact = p->learning_fact; h->pr.history_size = (int)p->history_size; h->pr.gamma_factor = p->gamma_factor;
    h->pr.pcg_tol = p->pcg_tol; h->pr.pcg_maxiters = p->pcg_maxiters;
    return 0;
}

// (Re)build the row image the ADMM kernel reads from the current CSR arrays (compact format).  `check`: decide whether the batch
// qualifies (every row <= 8 entries) -- once, at init; early fixing only ever removes entries from a row.
static int seg_build_rows(lpbox_seg_batch *h, int skip_done, bool check) {
    const bool off = getenv("LPBOX_SEG_NO_ROWIMG") != nullptr;      // experiments / tests: keep reading the CSR arrays
    if (!h->compact || off) return 0;
    if (check) { h->sv.use_ell = 1; SCK(cudaMemsetAsync(h->d_ell_flag.p, 0, sizeof(int), h->stream)); }
    if (!h->sv.use_ell) return 0;
    const dim3 grid((unsigned)h->B, (unsigned)std::max(1, std::min(64, (h->max_n + SEG_T - 1) / SEG_T)));
    seg_ell_build_kernel<<<grid, SEG_T, 0, h->stream>>>(h->sv, skip_done, h->d_ell_flag.p);
    SCK(cudaGetLastError());
    h->launches += 1;
    if (check) {
        int flag = 0;
        SCK(cudaMemcpyAsync(&flag, h->d_ell_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        SCK(cudaStreamSynchronize(h->stream));
        if (flag) h->sv.use_ell = 0;
    }
    return 0;
}

extern "C" int lpbox_seg_init(lpbox_seg_batch *h, const double *x0_all) {
    if (!h) return LPBOX_E_INVALID;
    SCK(cudaSetDevice(h->device));
    if (x0_all) {
        std::vector<double> xs((size_t)h->off_n[h->B], 0.0);
        long long o = 0;
        for (int i = 0; i < h->B; ++i) { memcpy(xs.data() + h->off_n[i], x0_all + o, sizeof(double) * (size_t)h->n0[i]); o += h->n0[i]; }
        SCK(cudaMemcpyAsync(h->sv.x, xs.data(), sizeof(double) * xs.size(), cudaMemcpyHostToDevice, h->stream));
        SCK(cudaStreamSynchronize(h->stream));
    }
    SCK(cudaEventRecord(h->ev0, h->stream));
    if (h->compact) seg_setup_kernel<true><<<h->B, SEG_T, h->smem, h->stream>>>(h->sv, h->pr, x0_all ? 1 : 0);
    else seg_setup_kernel<false><<<h->B, SEG_T, h->smem, h->stream>>>(h->sv, h->pr, x0_all ? 1 : 0);
    SCK(cudaGetLastError());
    h->launches += 1;
    int rc = seg_build_rows(h, 0, true); if (rc) return rc;
    SCK(cudaEventRecord(h->ev1, h->stream));
    rc = seg_sync_states(h); if (rc) return rc;
    float ms = 0; SCK(cudaEventElapsedTime(&ms, h->ev0, h->ev1)); h->last_ms = ms;
    h->inited = true;
    return 0;
}

template <int T>
static void seg_launch_admm(lpbox_seg_batch *h, const SegLaunch &la) {
    if (h->compact && h->sv.use_ell) seg_admm_kernel<true, T, true><<<h->grid, T, h->smem_admm, h->stream>>>(h->sv, h->pr, la);
    else if (h->compact) seg_admm_kernel<true, T, false><<<h->grid, T, h->smem_admm, h->stream>>>(h->sv, h->pr, la);
    else seg_admm_kernel<false, T, false><<<h->grid, T, h->smem_admm, h->stream>>>(h->sv, h->pr, la);
}

static int seg_run(lpbox_seg_batch *h, int iter_start, int iter_end, int l2f, int skip_done) {
    SegLaunch la{};
    la.iter_start = iter_start; la.iter_end = iter_end; la.l2f = l2f; la.skip_done = skip_done; la.n_work = h->B; la.counter = h->d_counter.p;
    SCK(cudaMemsetAsync(h->d_counter.p, 0, sizeof(int), h->stream));
    SCK(cudaEventRecord(h->ev0, h->stream));
    switch (h->threads) {
        case 256: seg_launch_admm<256>(h, la); break;
        case 192: seg_launch_admm<192>(h, la); break;
        default:  seg_launch_admm<160>(h, la); break;
    }
    SCK(cudaGetLastError());
    h->launches += 1;
    SCK(cudaEventRecord(h->ev1, h->stream));
    int rc = seg_sync_states(h); if (rc) return rc;
    float ms = 0; SCK(cudaEventElapsedTime(&ms, h->ev0, h->ev1)); h->last_ms = ms;
    return 0;
}

// ADMM_bqp_unconstrained_legacy for every image: energy[i] = int(cur_obj + _c)  (SEG.cpp:1200-1380)
extern "C" int lpbox_seg_solve(lpbox_seg_batch *h, int32_t *energy) {
    if (!h || !h->inited) { lpbox_set_error("call lpbox_seg_init first"); return LPBOX_E_INVALID; }
    SCK(cudaSetDevice(h->device));
    int rc = seg_run(h, 0, h->pr.max_iters, 0, 0);
    if (rc) return rc;
    if (energy) for (int i = 0; i < h->B; ++i) energy[i] = (int32_t)(h->h_st[i].cur_obj + h->h_st[i].cconst);
    return (int)(h->h_st[0].cur_obj + h->h_st[0].cconst);
}

// ADMM_bqp_unconstrained_l2f(iter_start, iter_end, vec, num) for every image (SEG.cpp:917-1195)
extern "C" int lpbox_seg_iters_l2f(lpbox_seg_batch *h, int iter_start, int iter_end, const double *vec_all, const int32_t *num, int32_t *ret) {
    if (!h || !h->inited) { lpbox_set_error("call lpbox_seg_init first"); return LPBOX_E_INVALID; }
    SCK(cudaSetDevice(h->device));
    const int skip_done = h->B > 1 ? 1 : 0;
    bool any = false;
    if (num) for (int i = 0; i < h->B; ++i) if (num[i] != 0) any = true;
    if (any && !vec_all) { lpbox_set_error("vec_all is NULL but some num[i] != 0"); return LPBOX_E_INVALID; }
    if (any) {
        std::vector<long long> off_vec(h->B + 1, 0);
        for (int i = 0; i < h->B; ++i) off_vec[i + 1] = off_vec[i] + h->h_st[i].n;
        for (int i = 0; i < h->B; ++i) {
            if (num[i] == 0) continue;
            int c = 0;
            const double *v = vec_all + off_vec[i];
            for (int k = 0; k < h->h_st[i].n; ++k) if (v[k] == 1.0 || v[k] == 0.0) c++;
            if (c != num[i]) { lpbox_set_error("num[i] does not match the number of fixed entries in vec"); return LPBOX_E_INVALID; }
        }
        SCK(cudaMemcpyAsync(h->d_off_vec.p, off_vec.data(), sizeof(long long) * (h->B + 1), cudaMemcpyHostToDevice, h->stream));
        SCK(cudaMemcpyAsync(h->d_vec.p, vec_all, sizeof(double) * (size_t)off_vec[h->B], cudaMemcpyHostToDevice, h->stream));
        SCK(cudaMemcpyAsync(h->d_num.p, num, sizeof(int) * (size_t)h->B, cudaMemcpyHostToDevice, h->stream));
        h->h2d_bytes += (int64_t)(sizeof(double) * (size_t)off_vec[h->B]);
    } else {
        SCK(cudaMemsetAsync(h->d_num.p, 0, sizeof(int) * (size_t)h->B, h->stream));
    }
    if (h->compact) seg_fix_kernel<true><<<h->B, SEG_T, 0, h->stream>>>(h->sv, h->pr, h->d_vec.p, h->d_off_vec.p, h->d_num.p, skip_done);
    else seg_fix_kernel<false><<<h->B, SEG_T, 0, h->stream>>>(h->sv, h->pr, h->d_vec.p, h->d_off_vec.p, h->d_num.p, skip_done);
    SCK(cudaGetLastError());
    h->launches += 1;
    int rc = seg_build_rows(h, skip_done, false);
    if (rc) return rc;
    rc = seg_run(h, iter_start, iter_end, 1, skip_done);
    if (rc) return rc;
    if (ret) for (int i = 0; i < h->B; ++i) ret[i] = h->h_st[i].last_ret;
    return h->h_st[0].last_ret;
}

// get_x_iters_d(ws) (SEG.cpp:833-845): row-major (n_cur x ws); the reference keeps 10 columns (SEG.cpp:924)
extern "C" int lpbox_seg_get_x_iters(lpbox_seg_batch *h, int i, int ws, double *out) {
    if (!h || i < 0 || i >= h->B || !out || ws <= 0) return LPBOX_E_INVALID;
    const SegInst &s = h->h_st[i];
    const int rows = s.xit_rows, cols = std::min(std::min(s.xit_cols, ws), h->hist_cap);
    std::vector<double> buf((size_t)std::max(cols, 0) * s.n0);
    if (cols > 0) {
        h->d2h_bytes += (int64_t)(sizeof(double) * buf.size());
        SCK(cudaMemcpyAsync(buf.data(), h->sv.hist + h->off_hist[i], sizeof(double) * buf.size(), cudaMemcpyDeviceToHost, h->stream));
        SCK(cudaStreamSynchronize(h->stream));
    }
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < ws; ++c) out[(size_t)r * ws + c] = (c < cols) ? buf[(size_t)c * s.n0 + r] : 0.0;
    return rows;
}

extern "C" int lpbox_seg_size(const lpbox_seg_batch *h) { return h ? h->B : LPBOX_E_INVALID; }
#define SCHK(h, i) if (!(h) || (i) < 0 || (i) >= (h)->B) return LPBOX_E_INVALID
extern "C" int lpbox_seg_get_n(lpbox_seg_batch *h, int i) { SCHK(h, i); return h->h_st[i].n; }
extern "C" int lpbox_seg_get_org_n(lpbox_seg_batch *h, int i) { SCHK(h, i); return h->h_st[i].n0; }
extern "C" int lpbox_seg_get_iter(lpbox_seg_batch *h, int i) { SCHK(h, i); return h->h_st[i].iter; }
static int seg_d2h(lpbox_seg_batch *h, void *dst, const void *src, size_t bytes) {
    if (!bytes) return 0;
    h->d2h_bytes += (int64_t)bytes;
    SCK(cudaSetDevice(h->device));
    SCK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    SCK(cudaStreamSynchronize(h->stream));
    return 0;
}
extern "C" int lpbox_seg_get_state(lpbox_seg_batch *h, int i, double *x, double *y1, double *y2, double *z1, double *z2) {
    SCHK(h, i);
    const size_t nb = sizeof(double) * (size_t)h->h_st[i].n;
    const long long on = h->off_n[i];
    int rc = 0;
    if (x) rc |= seg_d2h(h, x, h->sv.x + on, nb);
    if (y1) rc |= seg_d2h(h, y1, h->sv.y1 + on, nb);
    if (y2) rc |= seg_d2h(h, y2, h->sv.y2 + on, nb);
    if (z1) rc |= seg_d2h(h, z1, h->sv.z1 + on, nb);
    if (z2) rc |= seg_d2h(h, z2, h->sv.z2 + on, nb);
    return rc ? LPBOX_E_CUDA : 0;
}
// get_x_sol (SEG.cpp:895-915): fixed values + 1[x >= 0.5] of the remaining variables, original indexing
extern "C" int lpbox_seg_get_x_sol(lpbox_seg_batch *h, int i, double *out) {
    SCHK(h, i);
    if (!out) return LPBOX_E_INVALID;
    const SegInst &s = h->h_st[i];
    const long long on = h->off_n[i];
    std::vector<double> x(s.n), rval(s.n_ret);
    std::vector<int> left(s.n), ridx(s.n_ret);
    if (seg_d2h(h, x.data(), h->sv.x + on, sizeof(double) * (size_t)s.n) || seg_d2h(h, left.data(), h->sv.left_idx + on, sizeof(int) * (size_t)s.n) ||
        seg_d2h(h, ridx.data(), h->sv.ret_idx + on, sizeof(int) * (size_t)s.n_ret) || seg_d2h(h, rval.data(), h->sv.ret_val + on, sizeof(double) * (size_t)s.n_ret))
        return LPBOX_E_CUDA;
    for (int q = 0; q < s.n_ret; ++q) out[ridx[q]] = rval[q];
    for (int q = 0; q < s.n; ++q) out[left[q]] = (x[q] >= 0.5) ? 1.0 : 0.0;
    return s.n0;
}
// get_final_obj (SEG.cpp:868-893): x'Ax + b'x of the assembled binary solution on the ORIGINAL problem, + _c
extern "C" double lpbox_seg_get_final_obj(lpbox_seg_batch *h, int i) {
    if (!h || i < 0 || i >= h->B) return NAN;
    const SegInst &s = h->h_st[i];
    std::vector<double> xs(s.n0, 0.0), ax(s.n0), b(s.n0), va(s.nnz0);
    std::vector<int> rp(s.n0 + 1), ci(s.nnz0);
    if (lpbox_seg_get_x_sol(h, i, xs.data()) < 0) return NAN;
    if (seg_d2h(h, rp.data(), h->d_rp_org.p + h->off_n[i] + 4 * (size_t)i, sizeof(int) * ((size_t)s.n0 + 1)) ||
        seg_fetch_matrix(h, i, h->d_ci_org.p, h->d_val_org.p, rp.data(), ci.data(), va.data()) ||
        seg_d2h(h, b.data(), h->d_b_org.p + h->off_n[i], sizeof(double) * (size_t)s.n0)) return NAN;
    for (int r = 0; r < s.n0; ++r) { double acc = 0.0; for (int k = rp[r]; k < rp[r + 1]; ++k) acc = acc + va[k] * xs[ci[k]]; ax[r] = acc; }
    std::vector<double> pr1(s.n0), pr2(s.n0);
    for (int r = 0; r < s.n0; ++r) { pr1[r] = xs[r] * ax[r]; pr2[r] = b[r] * xs[r]; }
    return (eigen_sum(pr1.data(), s.n0) + eigen_sum(pr2.data(), s.n0)) + s.cconst;
}
extern "C" int lpbox_seg_results(lpbox_seg_batch *h, lpbox_log_row *log) {
    if (!h || !log) return LPBOX_E_INVALID;
    for (int i = 0; i < h->B; ++i) {
        const SegInst &s = h->h_st[i];
        log[i].iters = (int32_t)s.admm_iters; log[i].status = s.status; log[i].cg_iters = s.cg_iters;
        log[i].obj = s.cur_obj + s.cconst; log[i].cur_bin_obj = s.cur_obj; log[i].n_left = s.n; log[i].infeasible = 0;
    }
    return 0;
}
extern "C" double lpbox_seg_last_kernel_ms(const lpbox_seg_batch *h) { return h ? h->last_ms : -1.0; }
extern "C" int64_t lpbox_seg_launch_count(const lpbox_seg_batch *h) { return h ? h->launches : -1; }
extern "C" int64_t lpbox_seg_h2d_bytes(const lpbox_seg_batch *h) { return h ? h->h2d_bytes : -1; }
extern "C" int64_t lpbox_seg_d2h_bytes(const lpbox_seg_batch *h) { return h ? h->d2h_bytes : -1; }
