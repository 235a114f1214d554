// Sparse adversarial attack: the Lp-Box ADMM update of the pixel mask G, batched over images.
//
// Reference: SparseAttack/SparseAttack/main_ori.py:626-743 (update_G), :502-623 (loop), :376-499 (update_G_l2f),
// utils.py:8-16.  fp32 like the reference.  The attacked classifier stays in PyTorch; everything else of an iteration
// is two fused kernels (one CTA per image, state never leaves the GPU, no `.item()` host syncs):
//   sa_pre_kernel : y1 (box), y2 (shifted lp-sphere), y3 (group-lasso prox over the segments), classifier input
//   sa_post_kernel: chain rule through clamp/normalise, grad_G, gradient step on G, dual updates z1..z4, history
// Element-wise expressions keep the reference's operation order (explicit __f*_rn, no FMA); reductions are deterministic
// trees (the reference's torch reductions have an implementation-defined order, so parity is to a tolerance).
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/lpbox_b200.h"

void lpbox_set_error(const std::string &s);

namespace {

constexpr int SA_T = 256;

__device__ __forceinline__ float fM(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fA(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fS(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fD(float a, float b) { return __fdiv_rn(a, b); }

// deterministic block sum (fixed tree), result broadcast to all threads
__device__ __forceinline__ float block_sum(float v, float *s_red) {
    for (int o = 16; o > 0; o >>= 1) v = fA(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = (lane < SA_T / 32) ? s_red[lane] : 0.0f;
    for (int o = 4; o > 0; o >>= 1) t = fA(t, __shfl_xor_sync(0xffffffffu, t, o));
    return __shfl_sync(0xffffffffu, t, 0);
}

struct SaPre {
    int n_elem, nseg, seg_stride, chan_elems, n_chan;
    const float *G, *z1, *z2, *z3, *images, *eps;
    const int32_t *seg_ptr, *seg_elems, *seg_of;     // per image when seg_stride != 0
    const float *mean, *stdv;                        // per channel
    float rho1, rho2, rho3, lambda2, minpix, maxpix, half_sqrt_n;
    float *y1, *y2, *y3, *image_s;
};

__global__ void __launch_bounds__(SA_T) sa_pre_kernel(SaPre a) {
    extern __shared__ float smem[];
    float *coef = smem;                 // [nseg]
    float *Cs = smem + a.nseg;          // [n_elem]  C = G + z3/rho3
    __shared__ float s_red[SA_T / 32];
    const int img = blockIdx.x, tid = threadIdx.x;
    const size_t o = (size_t)img * a.n_elem;
    const size_t so = (size_t)img * a.seg_stride, sp = a.seg_stride ? (size_t)img * (a.nseg + 1) : 0;
    // y1 (main_ori.py:652), shift for y2 (:653, utils.py:9), C (:656)
    float sq = 0.0f;
    for (int i = tid; i < a.n_elem; i += SA_T) {
        const float g = a.G[o + i];
        const float t1 = fA(g, fD(a.z1[o + i], a.rho1));
        a.y1[o + i] = fminf(fmaxf(t1, 0.0f), 1.0f);
        const float sh = fS(fA(g, fD(a.z2[o + i], a.rho2)), 0.5f);
        a.y2[o + i] = sh;
        sq = fA(sq, fM(sh, sh));
        Cs[i] = fA(g, fD(a.z3[o + i], a.rho3));
        // classifier input (:670-672): clamp(images + G*eps), then (x - mean) / std
        const int c = i / a.chan_elems;
        float im = fA(a.images[o + i], fM(g, a.eps[o + i]));
        im = fminf(fmaxf(im, a.minpix), a.maxpix);
        a.image_s[o + i] = fD(fS(im, a.mean[c]), a.stdv[c]);
    }
    const float norm2 = sqrtf(block_sum(sq, s_red));
    // group-lasso coefficients (:657-662): one thread per segment, sequential sum of squares over its elements
    for (int s = tid; s < a.nseg; s += SA_T) {
        float acc = 0.0f;
        for (int k = a.seg_ptr[sp + s]; k < a.seg_ptr[sp + s + 1]; ++k) { const float v = Cs[a.seg_elems[so + k]]; acc = fA(acc, fM(v, v)); }
        const float nrm = sqrtf(acc);
        coef[s] = fmaxf(fS(1.0f, fD(a.lambda2, fM(a.rho3, nrm))), 0.0f);
    }
    __syncthreads();
    for (int i = tid; i < a.n_elem; i += SA_T) {
        a.y2[o + i] = fA(fM(a.half_sqrt_n, fD(a.y2[o + i], norm2)), 0.5f);      // utils.py:15
        a.y3[o + i] = fM(coef[a.seg_of[so + i]], Cs[i]);                        // :663-664 (one non-zero group per element)
    }
}

struct SaPost {
    int n_elem, chan_elems;
    float *G, *z1, *z2, *z3, *z4;
    const float *y1, *y2, *y3, *grad_in, *images, *eps, *nw, *stdv;
    float lambda1, rho1, rho2, rho3, rho4, step, minpix, maxpix;
    const float *lambda1_img;   // per-image lambda1 (the batched lambda1 search), or NULL -> the scalar
    double rho4_d, k;
    float *hist;            // [n_img][n_elem] slot of this iteration, or NULL
};

__global__ void __launch_bounds__(SA_T) sa_post_kernel(SaPost a) {
    __shared__ float s_red[SA_T / 32];
    const int img = blockIdx.x, tid = threadIdx.x;
    const size_t o = (size_t)img * a.n_elem;
    float s = 0.0f;
    for (int i = tid; i < a.n_elem; i += SA_T) s = fA(s, a.G[o + i]);
    const float gsum = block_sum(s, s_red);                                      // G.sum().item()  (:700)
    const float c4 = (float)(a.rho4_d * ((double)gsum - a.k));                   // cur_rho4*(G.sum().item() - k), a python float
    const float z4 = a.z4[img];
    const float lambda1 = a.lambda1_img ? a.lambda1_img[img] : a.lambda1;
    float s2 = 0.0f;
    for (int i = tid; i < a.n_elem; i += SA_T) {
        const float g = a.G[o + i], e = a.eps[o + i], w = a.nw[o + i];
        // autograd through Normalization -> clamp -> mul (main_ori.py:670-672)
        const float pre = fA(a.images[o + i], fM(g, e));
        float cg = fD(a.grad_in[o + i], a.stdv[i / a.chan_elems]);
        cg = (pre >= a.minpix && pre <= a.maxpix) ? cg : 0.0f;
        cg = fM(cg, e);
        // grad_G (:697-700), left to right
        float gr = fM(fM(fM(fM(fM(2.0f, g), e), e), w), w);
        gr = fA(gr, fM(lambda1, cg));
        gr = fA(gr, a.z1[o + i]); gr = fA(gr, a.z2[o + i]); gr = fA(gr, a.z3[o + i]);
        gr = fA(gr, z4);
        gr = fA(gr, fM(a.rho1, fS(g, a.y1[o + i])));
        gr = fA(gr, fM(a.rho2, fS(g, a.y2[o + i])));
        gr = fA(gr, fM(a.rho3, fS(g, a.y3[o + i])));
        gr = fA(gr, c4);
        const float gn = fS(g, fM(a.step, gr));                                  // :702
        a.G[o + i] = gn;
        if (a.hist) a.hist[o + i] = gn;
        a.z1[o + i] = fA(a.z1[o + i], fM(a.rho1, fS(gn, a.y1[o + i])));          // :718-720
        a.z2[o + i] = fA(a.z2[o + i], fM(a.rho2, fS(gn, a.y2[o + i])));
        a.z3[o + i] = fA(a.z3[o + i], fM(a.rho3, fS(gn, a.y3[o + i])));
        s2 = fA(s2, gn);
    }
    const float gsum2 = block_sum(s2, s_red);
    if (tid == 0) a.z4[img] = fA(z4, (float)(a.rho4_d * ((double)gsum2 - a.k)));  // :721
}

// update_G_l2f's overwrite (main_ori.py:476-485): p > hi -> 1, p < lo -> 0, else the window's last iterate
__global__ void sa_apply_policy_kernel(long long n, const float *__restrict__ scores, const float *__restrict__ last, float hi, float lo,
                                       float *__restrict__ G, int *__restrict__ counts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float p = scores[i];
    float g = last[i];
    if (p > hi) { g = 1.0f; atomicAdd(&counts[0], 1); }
    else if (p < lo) { g = 0.0f; atomicAdd(&counts[1], 1); }
    G[i] = g;
}


// ---- the perturbation step either side of update_G: update_epsilon, main_ori.py:310-354 ----------------
// classifier input: (clamp(images + eps*G, minpix, maxpix) - mean) / std   (:317-319)
__global__ void __launch_bounds__(SA_T) sa_eps_pre_kernel(int n_elem, int chan_elems, const float *__restrict__ images, const float *__restrict__ eps,
                                                          const float *__restrict__ G, const float *__restrict__ mean, const float *__restrict__ stdv,
                                                          float minpix, float maxpix, float *__restrict__ image_s) {
    const size_t o = (size_t)blockIdx.x * n_elem;
    for (int i = threadIdx.x; i < n_elem; i += SA_T) {
        const int c = i / chan_elems;
        float im = fA(images[o + i], fM(eps[o + i], G[o + i]));
        im = fminf(fmaxf(im, minpix), maxpix);
        image_s[o + i] = fD(fS(im, mean[c]), stdv[c]);
    }
}
// eps <- eps - step * (2*eps*G*G*w*w + lambda1 * dLoss/deps)   (:341-343), dLoss/deps by the chain rule through
// Normalization -> clamp -> mul
__global__ void __launch_bounds__(SA_T) sa_eps_post_kernel(int n_elem, int chan_elems, float *__restrict__ eps, const float *__restrict__ G,
                                                           const float *__restrict__ grad_in, const float *__restrict__ images,
                                                           const float *__restrict__ nw, const float *__restrict__ stdv, float lambda1,
                                                           const float *__restrict__ lambda1_img, float step, float minpix, float maxpix) {
    const int img = blockIdx.x;
    const size_t o = (size_t)img * n_elem;
    const float lam = lambda1_img ? lambda1_img[img] : lambda1;
    for (int i = threadIdx.x; i < n_elem; i += SA_T) {
        const float e = eps[o + i], g = G[o + i], w = nw[o + i];
        const float pre = fA(images[o + i], fM(e, g));
        float cg = fD(grad_in[o + i], stdv[i / chan_elems]);
        cg = (pre >= minpix && pre <= maxpix) ? cg : 0.0f;
        cg = fM(cg, g);
        float gr = fM(fM(fM(fM(fM(2.0f, e), g), g), w), w);
        gr = fA(gr, fM(lam, cg));
        eps[o + i] = fS(e, fM(step, gr));
    }
}

__device__ __forceinline__ float block_max(float v, float *s_red) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = (lane < SA_T / 32) ? s_red[lane] : 0.0f;
    for (int o = 4; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
    return __shfl_sync(0xffffffffu, t, 0);
}
// compute_statistics (utils.py:77-96) per image: out[img] = {G_sum, L0, L1, L2, Li, WL1, WL2, WLi} of
// noise = clamp(images + eps*G) - images and noise * Weight; plus ||G*eps*w||_2^2 (the l2 term of compute_loss, utils.py:27)
__global__ void __launch_bounds__(SA_T) sa_stats_kernel(int n_elem, const float *__restrict__ images, const float *__restrict__ eps,
                                                        const float *__restrict__ G, const float *__restrict__ nw, float minpix, float maxpix,
                                                        float *__restrict__ out) {
    __shared__ float s_red[SA_T / 32];
    const int img = blockIdx.x;
    const size_t o = (size_t)img * n_elem;
    float gs = 0.f, l0 = 0.f, l1 = 0.f, l2 = 0.f, li = 0.f, w1 = 0.f, w2 = 0.f, wi = 0.f, q = 0.f;
    for (int i = threadIdx.x; i < n_elem; i += SA_T) {
        const float g = G[o + i], e = eps[o + i], w = nw[o + i], im = images[o + i];
        const float eg = fM(e, g);
        const float noise = fS(fminf(fmaxf(fA(im, eg), minpix), maxpix), im);
        const float wn = fM(noise, w);
        gs = fA(gs, g);
        l0 = fA(l0, g > 0.5f ? 1.0f : 0.0f);
        l1 = fA(l1, fabsf(noise)); l2 = fA(l2, fM(noise, noise)); li = fmaxf(li, fabsf(noise));
        w1 = fA(w1, fabsf(wn)); w2 = fA(w2, fM(wn, wn)); wi = fmaxf(wi, fabsf(wn));
        const float t = fM(fM(g, e), w);
        q = fA(q, fM(t, t));
    }
    gs = block_sum(gs, s_red); l0 = block_sum(l0, s_red); l1 = block_sum(l1, s_red); l2 = block_sum(l2, s_red); li = block_max(li, s_red);
    w1 = block_sum(w1, s_red); w2 = block_sum(w2, s_red); wi = block_max(wi, s_red); q = block_sum(q, s_red);
    if (threadIdx.x == 0) {
        float *r = out + (size_t)img * 9;
        r[0] = gs; r[1] = l0; r[2] = l1; r[3] = sqrtf(l2); r[4] = li; r[5] = w1; r[6] = sqrtf(w2); r[7] = wi; r[8] = q;
    }
}

}  // namespace

#define SACK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { lpbox_set_error(std::string(#call) + ": " + cudaGetErrorString(e_)); return LPBOX_E_CUDA; } } while (0)

extern "C" int lpbox_sa_pre_dev(void *stream, int n_img, int n_elem, int n_chan, int nseg, int seg_per_image, const float *G, const float *z1,
                                const float *z2, const float *z3, const float *images, const float *eps, const int32_t *seg_ptr,
                                const int32_t *seg_elems, const int32_t *seg_of, const float *mean, const float *stdv, double rho1, double rho2,
                                double rho3, double lambda2, double minpix, double maxpix, float *y1, float *y2, float *y3, float *image_s) {
    if (n_img <= 0 || n_elem <= 0 || n_chan <= 0 || n_elem % n_chan || nseg <= 0) return LPBOX_E_INVALID;
    SaPre a;
    a.n_elem = n_elem; a.nseg = nseg; a.seg_stride = seg_per_image ? n_elem : 0; a.chan_elems = n_elem / n_chan; a.n_chan = n_chan;
    a.G = G; a.z1 = z1; a.z2 = z2; a.z3 = z3; a.images = images; a.eps = eps; a.seg_ptr = seg_ptr; a.seg_elems = seg_elems; a.seg_of = seg_of;
    a.mean = mean; a.stdv = stdv; a.rho1 = (float)rho1; a.rho2 = (float)rho2; a.rho3 = (float)rho3; a.lambda2 = (float)lambda2;
    a.minpix = (float)minpix; a.maxpix = (float)maxpix;
    a.half_sqrt_n = (float)(sqrt((double)n_elem) / 2);                           // (n ** (1/2)) / 2, a python float (utils.py:14-15)
    a.y1 = y1; a.y2 = y2; a.y3 = y3; a.image_s = image_s;
    const size_t smem = sizeof(float) * ((size_t)nseg + n_elem);
    if (smem > 48 * 1024) SACK(cudaFuncSetAttribute(sa_pre_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sa_pre_kernel<<<n_img, SA_T, smem, (cudaStream_t)stream>>>(a);
    SACK(cudaGetLastError());
    return 0;
}

extern "C" int lpbox_sa_post_dev(void *stream, int n_img, int n_elem, int n_chan, float *G, float *z1, float *z2, float *z3, float *z4,
                                 const float *y1, const float *y2, const float *y3, const float *grad_in, const float *images, const float *eps,
                                 const float *nw, const float *stdv, double lambda1, const float *lambda1_img, double rho1, double rho2,
                                 double rho3, double rho4, double step, double k, double minpix, double maxpix, float *hist_slot) {
    if (n_img <= 0 || n_elem <= 0 || n_chan <= 0 || n_elem % n_chan) return LPBOX_E_INVALID;
    SaPost a;
    a.n_elem = n_elem; a.chan_elems = n_elem / n_chan; a.G = G; a.z1 = z1; a.z2 = z2; a.z3 = z3; a.z4 = z4; a.y1 = y1; a.y2 = y2; a.y3 = y3;
    a.grad_in = grad_in; a.images = images; a.eps = eps; a.nw = nw; a.stdv = stdv; a.lambda1 = (float)lambda1; a.lambda1_img = lambda1_img; a.rho1 = (float)rho1;
    a.rho2 = (float)rho2; a.rho3 = (float)rho3; a.rho4 = (float)rho4; a.step = (float)step; a.minpix = (float)minpix; a.maxpix = (float)maxpix;
    a.rho4_d = rho4; a.k = k; a.hist = hist_slot;
    sa_post_kernel<<<n_img, SA_T, 0, (cudaStream_t)stream>>>(a);
    SACK(cudaGetLastError());
    return 0;
}

extern "C" int lpbox_sa_apply_policy_dev(void *stream, int64_t n, const float *scores, const float *last, double hi, double lo, float *G,
                                         int32_t *counts2) {
    if (n <= 0) return LPBOX_E_INVALID;
    SACK(cudaMemsetAsync(counts2, 0, 2 * sizeof(int32_t), (cudaStream_t)stream));
    sa_apply_policy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, scores, last, (float)hi, (float)lo, G, counts2);
    SACK(cudaGetLastError());
    return 0;
}

extern "C" int lpbox_sa_eps_pre_dev(void *stream, int n_img, int n_elem, int n_chan, const float *images, const float *eps, const float *G,
                                    const float *mean, const float *stdv, double minpix, double maxpix, float *image_s) {
    if (n_img <= 0 || n_elem <= 0 || n_chan <= 0 || n_elem % n_chan) return LPBOX_E_INVALID;
    sa_eps_pre_kernel<<<n_img, SA_T, 0, (cudaStream_t)stream>>>(n_elem, n_elem / n_chan, images, eps, G, mean, stdv, (float)minpix, (float)maxpix, image_s);
    SACK(cudaGetLastError());
    return 0;
}

extern "C" int lpbox_sa_eps_post_dev(void *stream, int n_img, int n_elem, int n_chan, float *eps, const float *G, const float *grad_in,
                                     const float *images, const float *nw, const float *stdv, double lambda1, const float *lambda1_img,
                                     double step, double minpix, double maxpix) {
    if (n_img <= 0 || n_elem <= 0 || n_chan <= 0 || n_elem % n_chan) return LPBOX_E_INVALID;
    sa_eps_post_kernel<<<n_img, SA_T, 0, (cudaStream_t)stream>>>(n_elem, n_elem / n_chan, eps, G, grad_in, images, nw, stdv, (float)lambda1, lambda1_img,
                                                                 (float)step, (float)minpix, (float)maxpix);
    SACK(cudaGetLastError());
    return 0;
}

extern "C" int lpbox_sa_stats_dev(void *stream, int n_img, int n_elem, const float *images, const float *eps, const float *G, const float *nw,
                                  double minpix, double maxpix, float *out9) {
    if (n_img <= 0 || n_elem <= 0) return LPBOX_E_INVALID;
    sa_stats_kernel<<<n_img, SA_T, 0, (cudaStream_t)stream>>>(n_elem, images, eps, G, nw, (float)minpix, (float)maxpix, out9);
    SACK(cudaGetLastError());
    return 0;
}
