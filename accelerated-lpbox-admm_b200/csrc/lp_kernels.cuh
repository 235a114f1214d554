// Persistent on-chip Lp-Box ADMM kernels for the inequality-constrained LP form (sm_100a).
//
// One CTA owns one problem instance for a whole window of ADMM iterations:
//   * the sparsity pattern of E (both orientations, uint16) is staged into shared memory with one 1-D TMA bulk copy
//     (cp.async.bulk + mbarrier),
//   * x, y1, y2, z1, z2, r, p, 1/diag live in REGISTERS of the thread that owns the element (EPT elements / thread),
//     y3, z4, f in registers of the thread that owns the constraint row,
//   * only vectors that other threads gather from (x or p for E v, E v for E^T, the reduction operands) go through
//     shared memory,
//   * projections, rhs assembly, the whole PCG solve, dual / rho updates and both stop tests are fused: nothing
//     touches HBM between iterations except the optional iterate history.
//
// PARITY MODE (the only mode in this file): every floating-point operation is issued in the order of the
// reference's compiled Eigen code: no FMA (explicit __dmul_rn/__dadd_rn), SpMV with one sequential
// accumulator per output row in ascending inner index, reductions as Eigen's SSE2 redux (four interleaved sequential
// chains).  Reference lines are cited at each step (LP.cpp = LinerProgramming/.../cython_solver/LPboxADMMsolver.cpp).
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "lp_types.h"

namespace lpb {

typedef uint16_t u16;

#define LPB_FOR_E _Pragma("unroll") for (int e = 0; e < EPT; ++e)

__device__ __forceinline__ double dM(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dA(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dS(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dD(double a, double b) { return __ddiv_rn(a, b); }

// ---- mbarrier / TMA bulk copy helpers (PTX) -----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared (TMA engine; SASS UBLKCP).  dst/src 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- Eigen-order reduction ----------------------------------------------------------------------------------------
// Executed by ONE warp.  Up to 8 independent reductions run side by side: lane l works on reduction q = l/4 as chain
// k = l%4 (Eigen's packet lanes: packet0 = {chain0, chain1}, packet1 = {chain2, chain3}).  Restates Eigen's
// redux_impl<..., LinearVectorizedTraversal, NoUnrolling> for Packet2d (call sites LP.cpp:277,288,300,306,311,323,425,
// 455,931-933).  Reduction q sums the products a_q[i] * c_q[i] formed on the fly (dot / squaredNorm expressions); with
// `ind` the second operand is the indicator 1[c_q[i] >= 0.5] (LP.cpp:1001-1005).  The per-lane operand pointers are
// set by the caller.  Result of reduction q is returned in every lane of its group.
__device__ __forceinline__ double prod_at(const double *a, const double *c, bool ind, int i) {
    double cv = c[i];
    if (ind) cv = (cv >= 0.5) ? 1.0 : 0.0;
    return dM(a[i], cv);
}
// one-operand variant of warp_redux_eigen2: each lane group sums its own array `v` of MATERIALISED terms (the owner threads
// stage the products / squares -- the same __dmul_rn the reduction would issue -- so the single reduction warp, which is the
// critical path of every CG iteration, only loads and adds)
__device__ __forceinline__ double warp_redux_eigen1(const double *v, int n) {
    const int lane = threadIdx.x & 31;
    const int k = lane & 3;
    const int a2 = n & ~3, a1 = n & ~1;
    double res;
    if (a1 > 2) {
        // chain k adds v[k], v[k+4], ... in order.  The loads of the NEXT four terms are issued before the four dependent adds
        // (loop-carried prefetch), otherwise ptxas, short of registers, serialises load -> add -> load on one register and
        // every step pays the shared-memory latency on top of the fp64 add latency.
        double acc = v[k];
        const double *pv = v + 4 + k;
        int left = a2 / 4 - 1;                         // terms still to add
        if (left >= 4) {
            double t0 = pv[0], t1 = pv[4], t2 = pv[8], t3 = pv[12];
            pv += 16; left -= 4;
            // two batches per trip, ping-pong between the t and u registers (no register moves in the loop)
#pragma unroll 1
            while (left >= 8) {
                const double u0 = pv[0], u1 = pv[4], u2 = pv[8], u3 = pv[12];
                acc = dA(acc, t0); acc = dA(acc, t1); acc = dA(acc, t2); acc = dA(acc, t3);
                t0 = pv[16]; t1 = pv[20]; t2 = pv[24]; t3 = pv[28];
                acc = dA(acc, u0); acc = dA(acc, u1); acc = dA(acc, u2); acc = dA(acc, u3);
                pv += 32; left -= 8;
            }
            if (left >= 4) {
                const double u0 = pv[0], u1 = pv[4], u2 = pv[8], u3 = pv[12];
                acc = dA(acc, t0); acc = dA(acc, t1); acc = dA(acc, t2); acc = dA(acc, t3);
                t0 = u0; t1 = u1; t2 = u2; t3 = u3;
                pv += 16; left -= 4;
            }
            acc = dA(acc, t0); acc = dA(acc, t1); acc = dA(acc, t2); acc = dA(acc, t3);
        }
        for (; left > 0; --left, pv += 4) acc = dA(acc, pv[0]);
        double hi = __shfl_down_sync(0xffffffffu, acc, 2);
        double l = dA(acc, hi);
        if (a1 > a2 && k < 2) l = dA(l, v[a2 + k]);
        double l1 = __shfl_down_sync(0xffffffffu, l, 1);
        res = dA(l, l1);
    } else if (a1 == 2) {
        res = dA(v[0], v[1]);
    } else {
        res = (n > 0) ? v[0] : 0.0;
    }
    if ((n & 1) && n > 1) res = dA(res, v[n - 1]);
    return __shfl_sync(0xffffffffu, res, lane & ~3);
}
__device__ __forceinline__ double warp_redux_eigen2(const double *__restrict__ a, const double *__restrict__ c, bool ind, int n) {
    const int lane = threadIdx.x & 31;
    const int k = lane & 3;
    const int a2 = n & ~3, a1 = n & ~1;
    double res;
    if (a1 > 2) {
        double acc = prod_at(a, c, ind, k);
        int i = 4 + k;
        // software-pipelined: loads and products are independent, only the adds form the chain
        for (; i + 28 < a2; i += 32) {
            double t0 = prod_at(a, c, ind, i), t1 = prod_at(a, c, ind, i + 4), t2 = prod_at(a, c, ind, i + 8),
                   t3 = prod_at(a, c, ind, i + 12), t4 = prod_at(a, c, ind, i + 16), t5 = prod_at(a, c, ind, i + 20),
                   t6 = prod_at(a, c, ind, i + 24), t7 = prod_at(a, c, ind, i + 28);
            acc = dA(acc, t0); acc = dA(acc, t1); acc = dA(acc, t2); acc = dA(acc, t3);
            acc = dA(acc, t4); acc = dA(acc, t5); acc = dA(acc, t6); acc = dA(acc, t7);
        }
        for (; i < a2; i += 4) acc = dA(acc, prod_at(a, c, ind, i));
        // packet_res0 = packet_res0 + packet_res1 : chain0+chain2 , chain1+chain3
        double hi = __shfl_down_sync(0xffffffffu, acc, 2);
        double l = dA(acc, hi);                                          // valid in k = 0,1
        if (a1 > a2 && k < 2) l = dA(l, prod_at(a, c, ind, a2 + k));     // one more packet
        double l1 = __shfl_down_sync(0xffffffffu, l, 1);
        res = dA(l, l1);                                                 // predux: lane0 + lane1 (valid in k = 0)
    } else if (a1 == 2) {
        res = dA(prod_at(a, c, ind, 0), prod_at(a, c, ind, 1));
    } else {
        res = (n > 0) ? prod_at(a, c, ind, 0) : 0.0;                     // n == 1 (coeff(0)); n == 0 -> 0
    }
    if ((n & 1) && n > 1) res = dA(res, prod_at(a, c, ind, n - 1));      // scalar tail (valid in k = 0)
    return __shfl_sync(0xffffffffu, res, lane & ~3);                     // broadcast chain-0 lane's value to its group
}
// single-array sum (used by the early-fix kernel on materialised products)
template <int R>
__device__ __forceinline__ double warp_redux_eigen(const double *red, int stride, int n) {
    const int lane = threadIdx.x & 31;
    const int q = lane >> 2, k = lane & 3;
    const double *v = red + (q < R ? q : 0) * stride;
    const int a2 = n & ~3, a1 = n & ~1;
    double res;
    if (a1 > 2) {
        double acc = v[k];
        for (int i = 4 + k; i < a2; i += 4) acc = dA(acc, v[i]);
        double hi = __shfl_down_sync(0xffffffffu, acc, 2);
        double l = dA(acc, hi);
        if (a1 > a2 && k < 2) l = dA(l, v[a2 + k]);
        double l1 = __shfl_down_sync(0xffffffffu, l, 1);
        res = dA(l, l1);
    } else if (a1 == 2) {
        res = dA(v[0], v[1]);
    } else {
        res = (n > 0) ? v[0] : 0.0;
    }
    if ((n & 1) && n > 1) res = dA(res, v[n - 1]);
    return __shfl_sync(0xffffffffu, res, lane & ~3);
}

// ---- sequential sparse dot products on the sliced-ELL image ---------------------------------------------------------
// out[perm[s]] = ((0 + c_1 v[i_1]) + c_2 v[i_2]) + ...  over the stored entries of slot s in ascending inner index
// (the reference's per-row order, SURVEY.md 8c rule 2).  Coefficient: COEF 0 -> 1 (unit E), 1 -> scale (unit rho4 E^T),
// 2 -> val[pos] (ELL order).  Index loads are warp-coalesced (32 consecutive uint16) and software-pipelined one batch
// ahead; the gathers of a batch are independent; only the adds form the dependent chain.
template <int T, int COEF>
__device__ __forceinline__ void seq_spmv_ell(const u16 *__restrict__ len, const u16 *__restrict__ sptr,
                                             const u16 *__restrict__ perm, const u16 *__restrict__ idx,
                                             const double *__restrict__ val, double scale, const double *__restrict__ v,
                                             int count, double *__restrict__ out) {
#pragma unroll 1
    for (int s = threadIdx.x; s < count; s += T) {
        const int L = len[s];
        int pos = (int)sptr[s >> 5] * 32 + (s & 31);     // entry k lives at pos + 32 k
        double acc = 0.0;
        int k = 0;
        if (L >= 4) {
            int i0 = idx[pos], i1 = idx[pos + 32], i2 = idx[pos + 64], i3 = idx[pos + 96];
#pragma unroll 1
            for (;;) {
                double t0 = v[i0], t1 = v[i1], t2 = v[i2], t3 = v[i3];
                if (COEF == 2) { t0 = dM(val[pos], t0); t1 = dM(val[pos + 32], t1); t2 = dM(val[pos + 64], t2); t3 = dM(val[pos + 96], t3); }
                k += 4; pos += 128;
                const bool more = (k + 4 <= L);
                if (more) { i0 = idx[pos]; i1 = idx[pos + 32]; i2 = idx[pos + 64]; i3 = idx[pos + 96]; }
                if (COEF == 1) { t0 = dM(scale, t0); t1 = dM(scale, t1); t2 = dM(scale, t2); t3 = dM(scale, t3); }
                acc = dA(acc, t0); acc = dA(acc, t1); acc = dA(acc, t2); acc = dA(acc, t3);
                if (!more) break;
            }
        }
#pragma unroll 1
        for (; k < L; ++k, pos += 32) {
            double t = v[idx[pos]];
            if (COEF == 1) t = dM(scale, t);
            if (COEF == 2) t = dM(val[pos], t);
            acc = dA(acc, t);
        }
        out[perm[s]] = acc;
    }
}

// Builds one orientation of the sliced-ELL image from the compressed one.  Block-cooperative (any blockDim, all
// threads must call).  ptr/idx(/val): compressed arrays (outer = row for the row image, column for the column image),
// perm: slot -> outer index.  s_w: shared scratch of >= count/32 + 2 ints.
__device__ __forceinline__ void build_ell(const u16 *ptr, const u16 *idx, const double *val, const u16 *perm, int count,
                                          u16 *o_len, u16 *o_sptr, u16 *o_idx, double *o_val, int *s_w) {
    const int ns = (count + 31) >> 5;
    for (int w = threadIdx.x; w <= ns; w += blockDim.x) s_w[w] = 0;
    __syncthreads();
    for (int s = threadIdx.x; s < count; s += blockDim.x) {
        const int o = perm[s];
        const int L = ptr[o + 1] - ptr[o];
        o_len[s] = (u16)L;
        atomicMax(&s_w[s >> 5], L);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int w = 0; w < ns; ++w) { int wd = s_w[w]; s_w[w] = run; o_sptr[w] = (u16)run; run += wd; }
        s_w[ns] = run; o_sptr[ns] = (u16)run;
    }
    __syncthreads();
    for (int s = threadIdx.x; s < count; s += blockDim.x) {
        const int o = perm[s];
        const int beg = ptr[o], L = ptr[o + 1] - beg;
        int pos = s_w[s >> 5] * 32 + (s & 31);
        for (int k = 0; k < L; ++k, pos += 32) { o_idx[pos] = idx[beg + k]; if (o_val) o_val[pos] = val[beg + k]; }
    }
    __syncthreads();
}

// shared-memory carve-up -----------------------------------------------------------------------------------------------
// Three n-vectors and two m-vectors are enough: every reduction takes its operands from vectors that are staged anyway
// (products are formed inside the reduction), and column-product results overwrite operands that are dead by then.
struct Smem {
    double *gv;      // [np]   vector being gathered by E v (x or p); second reduction operand (z); E^T z4 result
    double *a1;      // [np]   y2 pre-image / R4ET(f-y3) result / rhs / r
    double *a2;      // [np]   f - y3 (m entries) / column-product result / tmp = M p / x - y2
    double *t1;      // [mp]   E v
    double *wb;      // [mp]   copy of z4 gathered by E^T z4
    double *sc;      // [8]    reduction results / broadcast scalars
    double *ring;    // [16]   tail of obj_list
    double *ev_r, *ev_c, *r4v;    // ELL-order values: [evr_elems], [evc_elems], [evc_elems] (non-unit only)
    unsigned char *pat;
    uint64_t *bar;
};
__host__ __device__ inline size_t smem_bytes(int np, int mp, int pat_bytes, int evr_elems, int evc_elems) {
    // wb (a copy of z4, m entries) shares the n-sized a2 buffer with f - y3 (also m entries) whenever both fit: 2 mp <= np
    size_t d = (size_t)np * 3 + (size_t)mp * (2 * mp <= np ? 1 : 2) + 8 + 16 + (size_t)evr_elems + 2 * (size_t)evc_elems;
    return d * sizeof(double) + (size_t)pat_bytes + 16;
}
__device__ __forceinline__ Smem carve(unsigned char *base, int np, int mp, int pat_bytes, int evr_elems, int evc_elems) {
    Smem s;
    double *d = reinterpret_cast<double *>(base);
    s.gv = d; d += np;
    s.a1 = d; d += np;
    s.a2 = d; d += np;
    s.t1 = d; d += mp;
    if (2 * mp <= np) s.wb = s.a2 + mp; else { s.wb = d; d += mp; }
    s.sc = d; d += 8;
    s.ring = d; d += 16;
    s.ev_r = d; d += evr_elems;
    s.ev_c = d; d += evc_elems;
    s.r4v = d; d += evc_elems;
    s.pat = reinterpret_cast<unsigned char *>(d);
    s.bar = reinterpret_cast<uint64_t *>(s.pat + pat_bytes);
    return s;
}

// std_dev (LP.cpp:358-377) / compute_std_obj (:459-469) over the tail of obj_list kept in `ring`, with `obj` as the
// value about to be pushed (obj_len counts entries BEFORE the push).  pow(., 1/2) -> IEEE sqrt.
__device__ __forceinline__ double std_obj_after_push(const double *ring, long long obj_len, double obj, int history) {
    long long s = obj_len + 1;
    long long begin = (s <= history) ? 0 : s - history;
    int size = (int)(s - begin);
    double mean = 0.0;
    for (int i = 0; i < size; ++i) {
        long long idx = begin + i;
        double v = (idx == s - 1) ? obj : ring[idx & 15];
        mean = dA(mean, v);
    }
    mean = dD(mean, (double)size);
    double sd = 0.0;
    for (int i = 0; i < size; ++i) {
        long long idx = begin + i;
        double v = (idx == s - 1) ? obj : ring[idx & 15];
        double d = dS(v, mean);
        sd = dA(sd, dM(d, d));
    }
    sd = dD(sd, (double)(size - 1));
    double r = (sd == 0.0) ? 0.0 : sqrt(sd);
    return dD(r, fabs(obj));
}

// =====================================================================================================================
// The window kernel.  T threads, EPT elements (and rows) per thread: max(n, m) <= T*EPT.
// UNIT: all stored values of E are 1.0 (pattern-only matrices; rho4*E^T is one scalar).
// =====================================================================================================================
#ifndef LPB_CTAS_128
#define LPB_CTAS_128 7
#endif
template <int T, int EPT, bool UNIT>
__global__ void __launch_bounds__(T, (T <= 128 ? LPB_CTAS_128 : (T <= 256 ? 3 : 1)))
lp_admm_window_kernel(BatchView bv, Params pr, Launch la) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem S = carve(smem_raw, la.np, la.mp, la.pat_bytes, la.evr_elems, la.evc_elems);
    __shared__ int s_work;
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    constexpr int RW = T / 32 - 1;  // the warp that runs the reductions (its SpMV slots hold the shortest rows)
    constexpr int CE = UNIT ? 0 : 2;   // coefficient mode of products with E / E^T
    constexpr int CR = UNIT ? 1 : 2;   // coefficient mode of products with rho4 E^T
    const int np = la.np;
    uint32_t tma_phase = 0;

    if (tid == 0) { mbar_init(S.bar, 1); fence_mbar_init(); }
    __syncthreads();

    for (;;) {
        if (tid == 0) s_work = atomicAdd(la.counter, 1);
        __syncthreads();
        const int w = s_work;
        __syncthreads();
        if (w >= la.n_work) break;
        const int inst = la.work ? la.work[w] : w;
        InstState *stp = bv.st + inst;
        if ((la.skip_done && stp->done) || stp->n == 0) continue;   // uniform per CTA

        // ---------------- stage the instance --------------------------------------------------------------------
        const int n = stp->n, m = stp->m;
        const EllLayout PL = ell_layout(stp->n0, stp->m0, stp->rcap, stp->ccap);
        // an image larger than the shared-memory budget of this launch keeps its last array (the column indices) in global memory
        const bool spill = PL.bytes > la.pat_bytes;
        if (tid == 0) {
            fence_proxy_async();  // order earlier generic-proxy reads of the previous instance's blob before the overwrite
            mbar_expect_tx(S.bar, (uint32_t)(spill ? PL.o_cidx : PL.bytes));
            tma_load_1d(S.pat, bv.pat + bv.off_pat[inst], (uint32_t)(spill ? PL.o_cidx : PL.bytes), S.bar);
        }
        const long long on = bv.off_n[inst], om = bv.off_m[inst];
        const double *__restrict__ gb = bv.b + on;    // read-only during a window
        const double *__restrict__ gf = bv.f + om;
        double x[EPT], y1[EPT], y2[EPT], z1[EPT], z2[EPT], r[EPT], p[EPT], invd[EPT];
        double y3[EPT], z4[EPT];
        LPB_FOR_E {
            int j = tid + e * T;
            bool in = j < n;
            x[e] = in ? bv.x[on + j] : 0.0;   y1[e] = in ? bv.y1[on + j] : 0.0; y2[e] = in ? bv.y2[on + j] : 0.0;
            z1[e] = in ? bv.z1[on + j] : 0.0; z2[e] = in ? bv.z2[on + j] : 0.0;
            double pd = in ? bv.Pd[on + j] : 1.0;
            invd[e] = (pd != 0.0) ? dD(1.0, pd) : 1.0;   // value in use when rhoUpdated == 0 (Eigen: zero diagonal -> 1)
            r[e] = 0.0; p[e] = 0.0;
            bool rin = j < m;
            y3[e] = rin ? bv.y3[om + j] : 0.0; z4[e] = rin ? bv.z4[om + j] : 0.0;
        }
        double rho1 = stp->rho1, rho2 = stp->rho2, rho4 = stp->rho4, prho1 = stp->prho1, prho2 = stp->prho2,
               prho4 = stp->prho4, gamma = stp->gamma, ratio = stp->ratio, D = stp->D, r4s = stp->r4s,
               std_obj = stp->std_obj, cur_obj = stp->cur_obj, best_bin_obj = stp->best_bin_obj;
        int rhoUpdated = stp->rhoUpdated;
        long long obj_len = stp->obj_len, cg_total = 0, admm_total = 0;
        if (tid < 16) S.ring[tid] = stp->obj_ring[tid];
        const int evr_used = UNIT ? 0 : 32 * stp->rcap, evc_used = UNIT ? 0 : 32 * stp->ccap;
        if (!UNIT) {
            const long long er = bv.off_evr[inst], ec = bv.off_evc[inst];
            for (int k = tid; k < evr_used; k += T) S.ev_r[k] = bv.ev_r[er + k];
            for (int k = tid; k < evc_used; k += T) { S.ev_c[k] = bv.ev_c[ec + k]; S.r4v[k] = bv.r4v[ec + k]; }
        }
        mbar_wait(S.bar, tma_phase); tma_phase ^= 1;
        __syncthreads();
        const u16 *rlen = reinterpret_cast<const u16 *>(S.pat + PL.o_rlen);
        const u16 *rsptr = reinterpret_cast<const u16 *>(S.pat + PL.o_rsptr);
        const u16 *rperm = reinterpret_cast<const u16 *>(S.pat + PL.o_rperm);
        const u16 *ridx = reinterpret_cast<const u16 *>(S.pat + PL.o_ridx);
        const u16 *clen = reinterpret_cast<const u16 *>(S.pat + PL.o_clen);
        const u16 *csptr = reinterpret_cast<const u16 *>(S.pat + PL.o_csptr);
        const u16 *cperm = reinterpret_cast<const u16 *>(S.pat + PL.o_cperm);
        const u16 *cidx = reinterpret_cast<const u16 *>(S.pat + PL.o_cidx);
        const u16 *gcidx = reinterpret_cast<const u16 *>(bv.pat + bv.off_pat[inst] + PL.o_cidx);   // used instead of cidx when the image is spilled
#define LPB_COL_SPMV(COEF, VAL, SCALE, VIN, VOUT)                                                                     \
    do {                                                                                                              \
        if (spill) seq_spmv_ell<T, COEF>(clen, csptr, cperm, gcidx, VAL, SCALE, VIN, n, VOUT);                        \
        else seq_spmv_ell<T, COEF>(clen, csptr, cperm, cidx, VAL, SCALE, VIN, n, VOUT);                               \
    } while (0)
        const double pow_n = bv.pow_tab[n];               // std::pow(n, 1.0/p), p = 2 (LP.cpp:427)
        const int lane = tid & 31, rq = lane >> 2;        // reduction group of this lane (reduction warp only)

        int status = RUNNING;
        int iter = la.iter_start;
        int cc = 0;
        const bool lp_plain = (!la.l2f) && pr.guard_first_iter;

        for (; iter < la.iter_end; ++iter) {
            // ---- y1 (LP.cpp:806-809), y2 pre-image (:815, :424), gather x -------------------------------------
            LPB_FOR_E {
                int j = tid + e * T;
                double t = dA(x[e], dD(z1[e], rho1));
                y1[e] = (t > 1.0) ? 1.0 : ((t < 0.0) ? 0.0 : t);
                y2[e] = dS(dA(x[e], dD(z2[e], rho2)), 0.5);
                if (j < n) { S.a1[j] = dM(y2[e], y2[e]); S.gv[j] = x[e]; }     // squares staged for ||y|| (:425)
            }
            __syncthreads();
            // ---- ||y|| (:425) on the reduction warp, E x (:825) on the row slots -------------------------------
            if (warp == RW) {
                double v = warp_redux_eigen1(S.a1, n);
                if (lane == 0) S.sc[0] = v;
            }
            // E x of this iteration's x was already formed for the z4 update of the previous iteration (same operands, same
            // order -> same bits) and is still in t1; only the first iteration of a window has to compute it
            if (iter == la.iter_start) seq_spmv_ell<T, CE>(rlen, rsptr, rperm, ridx, S.ev_r, 0.0, S.gv, m, S.t1);
            __syncthreads();
            {
                const double nrm = sqrt(S.sc[0]);
                const double den = dM(2.0, nrm);
                LPB_FOR_E y2[e] = dA(dD(dM(y2[e], pow_n), den), 0.5);       // :427
            }
            // ---- y3 (:826-827); f - y3 and z4 staged for the two column products of the rhs ----------------------
            LPB_FOR_E {
                int i = tid + e * T;
                if (i < m) {
                    const double fi = gf[i];
                    double t = dS(dS(fi, S.t1[i]), dD(z4[e], rho4));
                    y3[e] = (t < 0.0) ? 0.0 : t;
                    S.a2[i] = dS(fi, y3[e]); S.wb[i] = z4[e];
                }
            }
            // ---- operator patch after a rho step (:851-866) and preconditioner refresh (:883-890) --------------
            if (iter != 0 && rhoUpdated) {
                const double c12 = dM(ratio, dA(prho1, prho2));
                const double c4 = dM(ratio, prho4);
                D = dA(D, c12);
                LPB_FOR_E {
                    int j = tid + e * T;
                    if (j < n) {
                        double pd = bv.Pd[on + j];
                        pd = dA(pd, c12);
                        pd = dA(pd, dM(c4, bv.Esq[on + j]));
                        bv.Pd[on + j] = pd;
                    }
                }
                if (UNIT) r4s = dM(pr.learning_fact, r4s);
                else for (int k = tid; k < evc_used; k += T) S.r4v[k] = dM(pr.learning_fact, S.r4v[k]);
            }
            if (rhoUpdated) {
                LPB_FOR_E {
                    int j = tid + e * T;
                    double pd = (j < n) ? bv.Pd[on + j] : 1.0;
                    invd[e] = (pd != 0.0) ? dD(1.0, pd) : 1.0;
                }
                rhoUpdated = 0;
            }
            __syncthreads();
            // ---- rhs (:872-878): R4ET (f - y3) -> a1, ET z4 -> gv ------------------------------------------------
            LPB_COL_SPMV(CR, S.r4v, r4s, S.a2, S.a1);
            LPB_COL_SPMV(CE, S.ev_c, 0.0, S.wb, S.gv);
            __syncthreads();
            // ---- PCG (:251-335), warm start x = y1 (:892) ------------------------------------------------------
            double rhs[EPT], xc[EPT];
            LPB_FOR_E {
                int j = tid + e * T;
                if (j < n) {
                    double t = dS(dA(dM(rho1, y1[e]), dM(rho2, y2[e])), dA(dA(gb[j], z1[e]), z2[e]));
                    t = dA(t, S.a1[j]);
                    rhs[e] = dS(t, S.gv[j]);
                    xc[e] = y1[e];
                    S.gv[j] = xc[e]; S.a1[j] = dM(rhs[e], rhs[e]);
                } else { rhs[e] = 0.0; xc[e] = 0.0; }
            }
            __syncthreads();
            if (warp == RW) {
                double v = warp_redux_eigen1(S.a1, n);                     // rhs.squaredNorm() :277
                if (lane == 0) S.sc[0] = v;
            }
            seq_spmv_ell<T, CE>(rlen, rsptr, rperm, ridx, S.ev_r, 0.0, S.gv, m, S.t1);
            __syncthreads();
            const double rhsNorm2 = S.sc[0];
            LPB_COL_SPMV(CR, S.r4v, r4s, S.t1, S.a2);
            __syncthreads();
            LPB_FOR_E {
                int j = tid + e * T;
                if (j < n) {
                    double mv = dA(dA(0.0, dM(D, xc[e])), S.a2[j]);          // D v (+) R4ET (E v)   :115-162
                    r[e] = dS(rhs[e], mv);                                   // :273
                    p[e] = dM(invd[e], r[e]);                                // :297
                    S.a1[j] = dM(r[e], r[e]); S.gv[j] = p[e]; S.a2[j] = dM(r[e], p[e]);
                }
            }
            __syncthreads();
            if (warp == RW) {
                double v = warp_redux_eigen1(rq == 0 ? S.a1 : S.a2, n);                // r.r :288, r.p :300
                if (lane == 0) S.sc[1] = v;
                if (lane == 4) S.sc[2] = v;
            }
            __syncthreads();
            int cg_it = 0;
            bool cg_fail = false;
            if (rhsNorm2 == 0.0) {                                           // :279-284
                LPB_FOR_E xc[e] = 0.0;
            } else {
                double threshold = dM(dM(pr.pcg_tol, pr.pcg_tol), rhsNorm2); // :287
                if (!(threshold > DBL_MIN)) threshold = DBL_MIN;
                double r2 = S.sc[1];
                double absNew = S.sc[2];
                if (!(r2 < threshold)) {                                     // :290-295
                    while (cg_it < pr.pcg_maxiters) {                        // gv holds p here
                        seq_spmv_ell<T, CE>(rlen, rsptr, rperm, ridx, S.ev_r, 0.0, S.gv, m, S.t1);
                        __syncthreads();
                        LPB_COL_SPMV(CR, S.r4v, r4s, S.t1, S.a2);
                        __syncthreads();
                        double tmp[EPT];
                        LPB_FOR_E {
                            int j = tid + e * T;
                            if (j < n) {
                                tmp[e] = dA(dA(0.0, dM(D, p[e])), S.a2[j]);   // :304
                                S.a2[j] = dM(p[e], tmp[e]);
                            } else tmp[e] = 0.0;
                        }
                        __syncthreads();
                        if (warp == RW) {
                            double v = warp_redux_eigen1(S.a2, n);                // p.dot(tmp) :306
                            if (lane == 0) S.sc[0] = v;
                        }
                        __syncthreads();
                        const double alpha = dD(absNew, S.sc[0]);
                        if (pr.alpha_bailout && alpha < 0.0) { cg_fail = true; break; }  // :307
                        double zz[EPT];
                        LPB_FOR_E {
                            int j = tid + e * T;
                            xc[e] = dA(xc[e], dM(alpha, p[e]));               // :308
                            r[e] = dS(r[e], dM(alpha, tmp[e]));               // :310
                            zz[e] = dM(invd[e], r[e]);                        // :320
                            if (j < n) { S.a1[j] = dM(r[e], r[e]); S.gv[j] = dM(r[e], zz[e]); }
                        }
                        __syncthreads();
                        if (warp == RW) {
                            double v = warp_redux_eigen1(rq == 0 ? S.a1 : S.gv, n);                // r.r :311, r.z :323
                            if (lane == 0) S.sc[1] = v;
                            if (lane == 4) S.sc[2] = v;
                        }
                        __syncthreads();
                        r2 = S.sc[1];
                        if (r2 < threshold) { cg_it++; break; }              // :315-318
                        const double absOld = absNew;
                        absNew = S.sc[2];
                        const double beta = dD(absNew, absOld);              // :324
                        LPB_FOR_E {
                            int j = tid + e * T;
                            p[e] = dA(zz[e], dM(beta, p[e]));                 // :325
                            if (j < n) S.gv[j] = p[e];
                        }
                        cg_it++;
                        __syncthreads();
                    }
                }
            }
            cg_total += cg_it;
            if (cg_fail) {
                if (la.l2f) { status = STOP_CG; break; }                    // :1450-1454 (x_sol keeps its old value)
                // plain loop ignores the return value: x_sol holds the partially updated iterate (:894)
            }
            LPB_FOR_E x[e] = xc[e];
            admm_total += 1;
            // ---- iterate history (:1472-1475) ----------------------------------------------------------------
            if ((la.l2f || la.record) && bv.hist_cap > 0) {
                if (cc < bv.hist_cap) {
                    double *h = bv.hist + bv.off_hist[inst] + (long long)cc * stp->n0;
                    LPB_FOR_E { int j = tid + e * T; if (j < n) h[j] = x[e]; }
                }
                cc++;
            }
            // ---- duals (:917-924) and stop-test operands (:931-933, :972, :1001-1011) ---------------------------
            {
                const double g1 = dM(gamma, rho1), g2 = dM(gamma, rho2);
                LPB_FOR_E {
                    int j = tid + e * T;
                    const double d1 = dS(x[e], y1[e]), d2 = dS(x[e], y2[e]);
                    z1[e] = dA(z1[e], dM(g1, d1));
                    z2[e] = dA(z2[e], dM(g2, d2));
                    if (j < n) { S.gv[j] = x[e]; S.a1[j] = d1; S.a2[j] = d2; }
                }
            }
            __syncthreads();
            seq_spmv_ell<T, CE>(rlen, rsptr, rperm, ridx, S.ev_r, 0.0, S.gv, m, S.t1);
            if (warp == RW) {
                // x.x, (x-y1)^2, (x-y2)^2, b.x, b.1[x>=0.5]  (:931-933, :972, :1001-1005) side by side
                const double *pa = (rq == 0) ? S.gv : (rq == 1) ? S.a1 : (rq == 2) ? S.a2 : gb;
                const double *pc = (rq == 1) ? S.a1 : (rq == 2) ? S.a2 : S.gv;
                double v = warp_redux_eigen2(pa, pc, rq == 4, n);
                if ((lane & 3) == 0 && lane < 20) S.sc[rq] = v;
                __syncwarp();
                if (lane == 0) {
                    const double obj = S.sc[3];
                    double so = std_obj;
                    if (obj_len + 1 >= (long long)pr.history_size) so = std_obj_after_push(S.ring, obj_len, obj, pr.history_size);
                    S.sc[5] = so;
                }
            }
            __syncthreads();
            {
                const double g4 = dM(gamma, rho4);
                const bool assign = lp_plain && (iter == la.iter_start);    // :920-921
                LPB_FOR_E {
                    int i = tid + e * T;
                    if (i < m) {
                        double t = dM(g4, dS(dA(S.t1[i], y3[e]), gf[i]));
                        z4[e] = assign ? t : dA(z4[e], t);
                    }
                }
            }
            {
                double temp0 = sqrt(S.sc[0]);                                // :931
                if (!(temp0 > 2.2204e-16)) temp0 = 2.2204e-16;
                const double c1 = dD(sqrt(S.sc[1]), temp0), c2 = dD(sqrt(S.sc[2]), temp0);
                const bool guard = lp_plain ? (iter != la.iter_start) : true;
                if (c1 <= pr.stop_threshold && c2 <= pr.stop_threshold && guard) { status = STOP_Y; break; }  // :934 / :1504
            }
            if ((iter + 1) % pr.rho_change_step == 0) {                      // :951-970
                prho1 = rho1; prho2 = rho2; prho4 = rho4;
                rho1 = dM(pr.learning_fact, rho1); rho2 = dM(pr.learning_fact, rho2); rho4 = dM(pr.learning_fact, rho4);
                double g = dM(gamma, pr.gamma_factor);
                gamma = (g < 1.0) ? 1.0 : g;
                rhoUpdated = 1;
                ratio = dS(pr.learning_fact, 1.0);
            }
            {
                const double obj = S.sc[3];                                  // :972-973
                if (tid == RW * 32) S.ring[obj_len & 15] = obj;              // written and read by the same thread only
                obj_len++;
                std_obj = S.sc[5];                                           // :974-976 (unchanged while size < history)
                if (std_obj <= pr.std_threshold) { status = STOP_STD; break; }   // :977
            }
            cur_obj = S.sc[4];                                               // :1001-1005
            if (best_bin_obj >= cur_obj) best_bin_obj = cur_obj;             // :1006-1009
            // all reads of S.sc / gv / a1 / a2 / t1 of this iteration are complete before the next iteration's first barrier
        }

        // ---------------- write the instance back ---------------------------------------------------------------
        __syncthreads();
        LPB_FOR_E {
            int j = tid + e * T;
            if (j < n) {
                bv.x[on + j] = x[e]; bv.y1[on + j] = y1[e]; bv.y2[on + j] = y2[e];
                bv.z1[on + j] = z1[e]; bv.z2[on + j] = z2[e];
            }
            if (j < m) { bv.y3[om + j] = y3[e]; bv.z4[om + j] = z4[e]; }
        }
        if (!UNIT) {
            const long long ec = bv.off_evc[inst];
            for (int k = tid; k < evc_used; k += T) bv.r4v[ec + k] = S.r4v[k];
        }
        if (tid < 16) stp->obj_ring[tid] = S.ring[tid];
        if (tid == 0) {
            stp->rho1 = rho1; stp->rho2 = rho2; stp->rho4 = rho4; stp->prho1 = prho1; stp->prho2 = prho2; stp->prho4 = prho4;
            stp->gamma = gamma; stp->ratio = ratio; stp->D = D; stp->r4s = r4s; stp->std_obj = std_obj;
            stp->cur_obj = cur_obj; stp->best_bin_obj = best_bin_obj; stp->rhoUpdated = rhoUpdated;
            stp->obj_len = obj_len; stp->cg_iters += cg_total; stp->admm_iters += admm_total;
            stp->iter = iter; stp->status = status;
            int ret;
            if (la.l2f) ret = (status != RUNNING || stp->norm_small) ? 1 : 0;    // :1505, :1542, :1452, :1223
            else ret = (status == STOP_STD) ? 1 : 0;                             // :978
            stp->last_ret = ret;
            // the window driver stops calling once a call returned 1 (LP.trainer:521); the plain driver calls once
            stp->done = la.l2f ? ret : (status != RUNNING);
            if (la.l2f || la.record) { stp->xit_cols = cc; if (!la.l2f) stp->xit_rows = n; }
        }
        __syncthreads();
    }
}

// =====================================================================================================================
// Set-up kernel: ADMM_lp_iters_init (LP.cpp:489-763) and/or update_expression (LP.cpp:2289-2404) for every instance.
// One CTA per instance, operating in HBM.  mode bit 0: initialise the iterate state; bit 1: rebuild the operator.
// =====================================================================================================================
static __global__ void lp_setup_kernel(BatchView bv, Params pr, int mode, int use_x0) {
    const int inst = blockIdx.x;
    InstState *st = bv.st + inst;
    const int n = st->n, m = st->m;
    const CsrLayout PL = csr_layout(st->n0, st->m0, st->nnz0);
    const unsigned char *pat = bv.csr + bv.off_csr[inst];
    const u16 *rowptr = reinterpret_cast<const u16 *>(pat + PL.o_rowptr);
    const u16 *colptr = reinterpret_cast<const u16 *>(pat + PL.o_colptr);
    const u16 *colidx = reinterpret_cast<const u16 *>(pat + PL.o_colidx);
    const u16 *rowidx = reinterpret_cast<const u16 *>(pat + PL.o_rowidx);
    const long long on = bv.off_n[inst], om = bv.off_m[inst], ov = bv.off_val ? bv.off_val[inst] : 0;
    const bool unit = st->unit != 0;
    const int tid = threadIdx.x, T = blockDim.x;
    __shared__ double s_best;
    __shared__ int s_w[80];
    if (mode & 4) {   // (re)build the sliced-ELL image from the compressed one
        const EllLayout EL = ell_layout(st->n0, st->m0, st->rcap, st->ccap);
        unsigned char *ell = bv.pat + bv.off_pat[inst];
        build_ell(rowptr, colidx, unit ? nullptr : bv.val_r + ov, reinterpret_cast<const u16 *>(ell + EL.o_rperm), m,
                  reinterpret_cast<u16 *>(ell + EL.o_rlen), reinterpret_cast<u16 *>(ell + EL.o_rsptr),
                  reinterpret_cast<u16 *>(ell + EL.o_ridx), unit ? nullptr : bv.ev_r + bv.off_evr[inst], s_w);
        build_ell(colptr, rowidx, unit ? nullptr : bv.val_c + ov, reinterpret_cast<const u16 *>(ell + EL.o_cperm), n,
                  reinterpret_cast<u16 *>(ell + EL.o_clen), reinterpret_cast<u16 *>(ell + EL.o_csptr),
                  reinterpret_cast<u16 *>(ell + EL.o_cidx), unit ? nullptr : bv.ev_c + bv.off_evc[inst], s_w);
    }
    if (mode & 1) {
        for (int j = tid; j < n; j += T) {
            double x0 = use_x0 ? bv.x[on + j] : 1.0;                         // :583-586
            bv.x[on + j] = x0; bv.y1[on + j] = x0; bv.y2[on + j] = x0;       // :714-715
            bv.z1[on + j] = 0.0; bv.z2[on + j] = 0.0;
            bv.left_idx[on + j] = j;
        }
        __syncthreads();
        for (int i = tid; i < m; i += T) {                                   // y3 = f - E x  (:720), z4 = 0 (:648)
            double acc = 0.0;
            for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
                double v = bv.x[on + colidx[k]];
                acc = dA(acc, unit ? v : dM(bv.val_r[ov + k], v));
            }
            bv.y3[om + i] = dS(bv.f[om + i], acc);
            bv.z4[om + i] = 0.0;
        }
        // best_bin_obj = b.dot(x_sol) (:726) in Eigen order -- one warp
        if (tid < 32) {
            // products staged through the (not yet used) history-free y1 copy is avoided: recompute per lane chain
            const int lane = tid, k = lane & 3;
            const int a2 = n & ~3, a1 = n & ~1;
            double res = 0.0;
            if (lane < 4) {
                if (a1 > 2) {
                    double acc = dM(bv.b[on + k], bv.x[on + k]);
                    for (int i = 4 + k; i < a2; i += 4) acc = dA(acc, dM(bv.b[on + i], bv.x[on + i]));
                    double hi = __shfl_down_sync(0xfu, acc, 2);
                    double l = dA(acc, hi);
                    if (a1 > a2 && k < 2) l = dA(l, dM(bv.b[on + a2 + k], bv.x[on + a2 + k]));
                    double l1 = __shfl_down_sync(0xfu, l, 1);
                    res = dA(l, l1);
                } else if (a1 == 2) {
                    res = dA(dM(bv.b[on], bv.x[on]), dM(bv.b[on + 1], bv.x[on + 1]));
                } else if (n == 1) {
                    res = dM(bv.b[on], bv.x[on]);
                }
                if ((n & 1) && n > 1) res = dA(res, dM(bv.b[on + n - 1], bv.x[on + n - 1]));
                if (lane == 0) s_best = res;
            }
        }
        __syncthreads();
        if (tid == 0) {
            st->rho1 = st->rho2 = st->rho4 = pr.initial_rho;                 // :629-636
            st->prho1 = st->prho2 = st->prho4 = pr.initial_rho;
            st->gamma = pr.gamma_val; st->ratio = 0.0;
            st->rhoUpdated = 1; st->std_obj = 1.0; st->cur_obj = 0.0; st->best_bin_obj = s_best;
            st->sum_fix_obj = 0.0; st->fix_obj = 0.0; st->prev_obj = 0.0; st->prev_sum = 0.0;
            st->obj_len = 0; st->cg_iters = 0; st->admm_iters = 0; st->iter = 0; st->status = RUNNING; st->done = 0;
            st->last_ret = 0; st->n_ret = 0; st->fix_sum = 0; st->xit_rows = 0; st->xit_cols = 0; st->norm_small = 0;
            for (int k = 0; k < 16; ++k) st->obj_ring[k] = 0.0;
        }
        __syncthreads();
    }
    if (mode & 2) {
        const double rho1 = st->rho1, rho2 = st->rho2, rho4 = st->rho4;
        const double D = dA(0.0, dA(rho1, rho2));                            // :2339-2343
        for (int j = tid; j < n; j += T) {
            double e = 0.0;                                                  // :2379-2390
            for (int k = colptr[j]; k < colptr[j + 1]; ++k) {
                if (unit) e = dA(e, 1.0);
                else { double v = bv.val_c[ov + k]; if (v != 0.0) e = dA(e, dM(v, v)); }
            }
            bv.Esq[on + j] = e;
            bv.Pd[on + j] = dA(D, dM(rho4, e));                              // :2351, :2391
        }
        if (!unit) {                                                         // :2292-2293 (ELL order; padding is never read)
            const long long ec = bv.off_evc[inst];
            for (int k = tid; k < 32 * st->ccap; k += T) bv.r4v[ec + k] = dM(rho4, bv.ev_c[ec + k]);
        }
        __syncthreads();
        if (tid == 0) { st->D = D; st->r4s = dM(rho4, 1.0); }
    }
}

}  // namespace lpb
