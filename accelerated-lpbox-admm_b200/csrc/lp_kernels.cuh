// Persistent on-chip Lp-Box ADMM kernels for the inequality-constrained LP form (sm_100a).
//
// One CTA owns one problem instance for a whole window of ADMM iterations:
//   * the sparsity pattern of E (both orientations, padded sliced-ELL of uint16 operand offsets, lp_types.h) is staged
//     into shared memory with one 1-D TMA bulk copy (cp.async.bulk + mbarrier),
//   * a thread OWNS the columns (variables) and rows (constraints) of its SpMV slots: the column product of a variable
//     lands in the registers of the thread that keeps x, r, p, 1/diag of that variable, so E^T(.) results never pass
//     through shared memory; the gathered operand vectors (x or p -> G, E v -> T1) are written in slot order (coalesced),
//   * only x, r, p, 1/diag live in registers across the PCG loop; y1, y2, z1, z2, b, y3, z4, f -- touched once or twice per
//     ADMM iteration -- are parked per CTA in an L2-resident scratch (slot order, coalesced, same thread writes and reads),
//   * reduction operands are staged CHAIN-MAJOR (Eigen's four interleaved chains are four contiguous runs), so the
//     reduction warp reads 16 bytes per lane per load and two reductions side by side use every bank exactly once,
//   * the warp that runs the sequential reductions rotates with the arrival order of the CTA on its SM, so the reduction
//     warps of the resident CTAs sit on different warp schedulers,
//   * projections, rhs assembly, the whole PCG solve, dual / rho updates and both stop tests are fused: nothing
//     touches HBM between iterations except the optional iterate history.
//
// PARITY MODE (default): every floating-point operation is issued in the order of the reference's compiled Eigen code
// (SURVEY.md §8c): no FMA (explicit __dmul_rn/__dadd_rn), SpMV with one sequential accumulator per output row in
// ascending inner index, reductions as Eigen's SSE2 redux (four interleaved sequential chains).  Reference lines are
// cited at each step (LP.cpp = LinerProgramming/.../cython_solver/LPboxADMMsolver.cpp).
// FAST MODE (template flag, opt-in through lpbox_batch_set_mode): the same iteration with tree reductions over all
// warps (shuffles) and fused multiply-adds in the vector updates -- NOT bit-identical, never used for parity claims.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "lp_types.h"

namespace lpb {

typedef uint16_t u16;

#define LPB_FOR_E _Pragma("unroll") for (int e = 0; e < EPT; ++e)
// a value that is the same in every lane, handed to the compiler as such (REDUX writes a uniform register): loops and
// branches on it need no divergence bookkeeping
#define LPB_UNIFORM(x) ((int)__reduce_max_sync(0xffffffffu, (unsigned)(x)))

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
// explicit shared-window accesses (32-bit addresses): no generic-pointer conversions in the hot loops
__device__ __forceinline__ double lds64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ double2 lds128(uint32_t a) {
    double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v;
}
// (a 32-bit destination: ld.u16 zero-extends, so no separate widening instructions follow the load)
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds_u32x2(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u32x2(uint32_t a, uint2 v) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ unsigned smid() { unsigned v; asm volatile("mov.u32 %0, %%smid;" : "=r"(v)); return v; }

// ---- Eigen-order reductions on chain-major buffers -------------------------------------------------------------------
// Executed by ONE warp.  Up to 8 independent reductions run side by side: lane l works on reduction q = l/4 as chain
// k = l%4 (Eigen's packet lanes: packet0 = {chain0, chain1}, packet1 = {chain2, chain3}).  Restates Eigen's
// redux_impl<..., LinearVectorizedTraversal, NoUnrolling> for Packet2d (call sites LP.cpp:277,288,300,306,311,323,425,
// 455,931-933).  Element j of the vector lives at v[(j & 3) * CH + (j >> 2)], so chain k is the contiguous run
// v[k*CH .. k*CH + n/4): the lane loads two terms per 16-byte load.  Result of reduction q in every lane of its group.
//
// warp_redux_cm1: every lane group sums its own array `v` of MATERIALISED terms (the owner threads stage the products /
// squares -- the same __dmul_rn the reduction would issue -- so this warp, the critical path of every CG iteration, only
// loads and adds).
// `groups` = number of reductions running side by side: only their 4 * groups lanes issue loads (a 16-byte shared load is
// processed a quarter-warp at a time, so idle lanes replaying another group's addresses would cost extra wavefronts).  All
// branches that contain shuffles depend on n only, which is the same in every lane.
__device__ __forceinline__ double warp_redux_cm1(uint32_t v, int n, int CH, int groups) {
    const int lane = threadIdx.x & 31;
    const int k = lane & 3;
    const bool act = lane < 4 * groups;
    const int cnt = n >> 2;                  // terms per chain in the vectorised part (a2 / 4)
    const int a2 = n & ~3, a1 = n & ~1;
    const uint32_t pv = v + (uint32_t)(k * CH) * 8u;
    double res;
    if (a1 > 2) {
        double acc = 0.0;
        if (act) {
            // chain k adds pv[0], pv[1], ... in order.  The loads of the NEXT pairs are issued before the dependent adds
            // (loop-carried prefetch, two batches per trip with ping-pong registers -> no register moves in the loop).
            const int full = cnt >> 1;           // complete pairs
            double2 h = lds128(pv);
            acc = h.x;
            if (cnt > 1) acc = dA(acc, h.y);
            uint32_t q = pv + 16;                // next pair
            int left = full - 1;                 // complete pairs still to add
            if (left >= 4) {
                double2 t0 = lds128(q), t1 = lds128(q + 16), t2 = lds128(q + 32), t3 = lds128(q + 48);
                q += 64; left -= 4;
#pragma unroll 1
                while (left >= 8) {
                    const double2 u0 = lds128(q), u1 = lds128(q + 16), u2 = lds128(q + 32), u3 = lds128(q + 48);
                    acc = dA(acc, t0.x); acc = dA(acc, t0.y); acc = dA(acc, t1.x); acc = dA(acc, t1.y);
                    acc = dA(acc, t2.x); acc = dA(acc, t2.y); acc = dA(acc, t3.x); acc = dA(acc, t3.y);
                    t0 = lds128(q + 64); t1 = lds128(q + 80); t2 = lds128(q + 96); t3 = lds128(q + 112);
                    acc = dA(acc, u0.x); acc = dA(acc, u0.y); acc = dA(acc, u1.x); acc = dA(acc, u1.y);
                    acc = dA(acc, u2.x); acc = dA(acc, u2.y); acc = dA(acc, u3.x); acc = dA(acc, u3.y);
                    q += 128; left -= 8;
                }
                if (left >= 4) {
                    const double2 u0 = lds128(q), u1 = lds128(q + 16), u2 = lds128(q + 32), u3 = lds128(q + 48);
                    acc = dA(acc, t0.x); acc = dA(acc, t0.y); acc = dA(acc, t1.x); acc = dA(acc, t1.y);
                    acc = dA(acc, t2.x); acc = dA(acc, t2.y); acc = dA(acc, t3.x); acc = dA(acc, t3.y);
                    t0 = u0; t1 = u1; t2 = u2; t3 = u3;
                    q += 64; left -= 4;
                }
                acc = dA(acc, t0.x); acc = dA(acc, t0.y); acc = dA(acc, t1.x); acc = dA(acc, t1.y);
                acc = dA(acc, t2.x); acc = dA(acc, t2.y); acc = dA(acc, t3.x); acc = dA(acc, t3.y);
            }
            for (; left > 0; --left, q += 16) { const double2 t = lds128(q); acc = dA(acc, t.x); acc = dA(acc, t.y); }
            if ((cnt & 1) && cnt > 1) acc = dA(acc, lds64(pv + (uint32_t)(cnt - 1) * 8u));
        }
        // packet_res0 = packet_res0 + packet_res1 : chain0+chain2 , chain1+chain3
        double hi = __shfl_down_sync(0xffffffffu, acc, 2);
        double l = dA(acc, hi);                                          // valid in k = 0,1
        if (act && a1 > a2 && k < 2) l = dA(l, lds64(pv + (uint32_t)cnt * 8u)); // one more packet: elements a2, a2+1
        double l1 = __shfl_down_sync(0xffffffffu, l, 1);
        res = dA(l, l1);                                                 // predux: lane0 + lane1 (valid in k = 0)
    } else if (a1 == 2) {
        const double t = act ? lds64(pv) : 0.0;                          // element k (k = 0,1)
        const double t1 = __shfl_down_sync(0xffffffffu, t, 1);
        res = dA(t, t1);
    } else {
        res = (act && n > 0) ? lds64(pv) : 0.0;                          // n == 1 (coeff(0)); n == 0 -> 0
    }
    if (act && (n & 1) && n > 1) res = dA(res, lds64(v + (uint32_t)(((n - 1) & 3) * CH + ((n - 1) >> 2)) * 8u));   // scalar tail (valid in k = 0)
    return __shfl_sync(0xffffffffu, res, lane & ~3);                     // broadcast chain-0 lane's value to its group
}
// warp_redux_cm2: terms a[i] * c'[i] formed on the fly; mode 0: c' = c, mode 1: c' = 1[c >= 0.5] (LP.cpp:1001-1005),
// mode 2: c' = 1 (materialised terms in a; x * 1.0 is exact).  Used once per ADMM iteration (x.x, b.x, b.1[x>=.5] and the
// next iteration's ||y||^2 side by side from three buffers).
__device__ __forceinline__ double cm_term(double a, double c, int mode) {
    if (mode == 1) c = (c >= 0.5) ? 1.0 : 0.0;
    if (mode == 2) c = 1.0;
    return dM(a, c);
}
__device__ __forceinline__ double warp_redux_cm2(uint32_t a, uint32_t c, int mode, int n, int CH, int groups) {
    const int lane = threadIdx.x & 31;
    const int k = lane & 3;
    const bool act = lane < 4 * groups;
    const int cnt = n >> 2;
    const int a2 = n & ~3, a1 = n & ~1;
    const uint32_t pa = a + (uint32_t)(k * CH) * 8u, pc = c + (uint32_t)(k * CH) * 8u;
    double res;
    if (a1 > 2) {
        double acc = 0.0;
        if (act) {
            const int full = cnt >> 1;
            double2 ha = lds128(pa), hc = lds128(pc);
            acc = cm_term(ha.x, hc.x, mode);
            if (cnt > 1) acc = dA(acc, cm_term(ha.y, hc.y, mode));
            uint32_t o = 16;
            int left = full - 1;
            if (left >= 2) {
                double2 ta0 = lds128(pa + o), tc0 = lds128(pc + o), ta1 = lds128(pa + o + 16), tc1 = lds128(pc + o + 16);
                o += 32; left -= 2;
#pragma unroll 1
                while (left >= 2) {
                    const double2 ua0 = lds128(pa + o), uc0 = lds128(pc + o), ua1 = lds128(pa + o + 16), uc1 = lds128(pc + o + 16);
                    const double p0 = cm_term(ta0.x, tc0.x, mode), p1 = cm_term(ta0.y, tc0.y, mode),
                                 p2 = cm_term(ta1.x, tc1.x, mode), p3 = cm_term(ta1.y, tc1.y, mode);
                    acc = dA(acc, p0); acc = dA(acc, p1); acc = dA(acc, p2); acc = dA(acc, p3);
                    ta0 = ua0; tc0 = uc0; ta1 = ua1; tc1 = uc1;
                    o += 32; left -= 2;
                }
                acc = dA(acc, cm_term(ta0.x, tc0.x, mode)); acc = dA(acc, cm_term(ta0.y, tc0.y, mode));
                acc = dA(acc, cm_term(ta1.x, tc1.x, mode)); acc = dA(acc, cm_term(ta1.y, tc1.y, mode));
            }
            for (; left > 0; --left, o += 16) {
                const double2 ta = lds128(pa + o), tc = lds128(pc + o);
                acc = dA(acc, cm_term(ta.x, tc.x, mode)); acc = dA(acc, cm_term(ta.y, tc.y, mode));
            }
            if ((cnt & 1) && cnt > 1) acc = dA(acc, cm_term(lds64(pa + (uint32_t)(cnt - 1) * 8u), lds64(pc + (uint32_t)(cnt - 1) * 8u), mode));
        }
        double hi = __shfl_down_sync(0xffffffffu, acc, 2);
        double l = dA(acc, hi);
        if (act && a1 > a2 && k < 2) l = dA(l, cm_term(lds64(pa + (uint32_t)cnt * 8u), lds64(pc + (uint32_t)cnt * 8u), mode));
        double l1 = __shfl_down_sync(0xffffffffu, l, 1);
        res = dA(l, l1);
    } else if (a1 == 2) {
        const double t = act ? cm_term(lds64(pa), lds64(pc), mode) : 0.0;
        const double t1 = __shfl_down_sync(0xffffffffu, t, 1);
        res = dA(t, t1);
    } else {
        res = (act && n > 0) ? cm_term(lds64(pa), lds64(pc), mode) : 0.0;
    }
    if (act && (n & 1) && n > 1) {
        const uint32_t o = (uint32_t)(((n - 1) & 3) * CH + ((n - 1) >> 2)) * 8u;
        res = dA(res, cm_term(lds64(a + o), lds64(c + o), mode));
    }
    return __shfl_sync(0xffffffffu, res, lane & ~3);
}
// single-array sum over a NATURALLY ordered array (used by the early-fix kernel on materialised products)
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

// ---- fast mode: tree sum of K per-thread partials over the whole CTA (NOT the reference's order) ---------------------
// `part` is a [NW][K] shared buffer; the caller alternates between two such buffers so that one barrier per call is enough.
template <int K, int NW>
__device__ __forceinline__ void block_sum(double (&v)[K], double *part) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if (lane == 0) part[warp * K + k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double s = part[k];
#pragma unroll
        for (int w = 1; w < NW; ++w) s += part[w * K + k];
        v[k] = s;
    }
}

// ---- sequential sparse dot products on the padded sliced-ELL image -------------------------------------------------
// One slot: ((0 + c_1 v[i_1]) + c_2 v[i_2]) + ... over the W entries of its slice row, ascending inner index (the
// reference's per-row order, SURVEY.md 8c rule 2; the trailing padding adds +0.0).  `idx` holds byte offsets of the
// operands from `sb` (start of dynamic shared memory).  COEF 0 -> coefficient 1 (unit E; a scalar factor is applied to
// the operand vector once per entry of that vector instead of once per stored entry -- the same product, the same bits),
// 2 -> val[pos] (ELL order).  Index loads are warp-coalesced (32 consecutive uint16) and software-pipelined one batch
// ahead; the gathers of a batch are independent; only the adds form the dependent chain.  W is warp-uniform.
// `idx` entries are ABSOLUTE shared-window addresses of the operands (lp_types.h).  GIDX: the index array is read from global
// memory (gidx, element index pos) instead of shared memory (address ia + 2 pos) -- images above the shared-memory budget.
// Offsets are stored in PAIRS (lp_types.h: ell_pos): one 32-bit load per lane fetches the offsets of two consecutive entries, so a
// warp reads 128 bytes = one full shared-memory wavefront per two entries instead of one 64-byte wavefront per entry.
template <bool GIDX>
__device__ __forceinline__ uint32_t ld_pair(uint32_t ia, const u16 *__restrict__ gidx, int pos) {   // pos: even u16 index
    return GIDX ? *reinterpret_cast<const uint32_t *>(gidx + pos) : lds32(ia + 2u * (uint32_t)pos);
}
template <bool GIDX>
__device__ __forceinline__ uint32_t ld_idx(uint32_t ia, const u16 *__restrict__ gidx, int pos) {
    return GIDX ? (uint32_t)gidx[pos] : lds16(ia + 2u * (uint32_t)pos);
}
// base: u16 index of the slice (32 * sptr[w]); W: width of the slice (warp-uniform)
template <int COEF, bool GIDX>
__device__ __forceinline__ double ell_dot(uint32_t ia, const u16 *__restrict__ gidx, const double *__restrict__ val, int base, int lane, int W,
                                          uint32_t add) {
    double acc = 0.0;
    int k = 0, pp = base + 2 * lane;                 // this lane's pair of entries (k, k + 1)
    if (W >= 4) {
        uint32_t p0 = ld_pair<GIDX>(ia, gidx, pp), p1 = ld_pair<GIDX>(ia, gidx, pp + 64);
#pragma unroll 1
        for (;;) {
            double t0 = lds64((p0 & 0xffffu) + add), t1 = lds64((p0 >> 16) + add), t2 = lds64((p1 & 0xffffu) + add), t3 = lds64((p1 >> 16) + add);
            if (COEF == 2) { t0 = dM(val[pp], t0); t1 = dM(val[pp + 1], t1); t2 = dM(val[pp + 64], t2); t3 = dM(val[pp + 65], t3); }
            k += 4; pp += 128;
            // the next two pairs are fetched unconditionally (what follows the arrays is readable, lp_types.h): past the end of
            // the slice they are simply not used
            p0 = ld_pair<GIDX>(ia, gidx, pp); p1 = ld_pair<GIDX>(ia, gidx, pp + 64);
            acc = dA(acc, t0); acc = dA(acc, t1); acc = dA(acc, t2); acc = dA(acc, t3);
            if (k + 4 > W) break;
        }
    }
    if (W & 2) {                                     // W is warp-uniform: straight-line tail
        const uint32_t p0 = ld_pair<GIDX>(ia, gidx, pp);
        double t0 = lds64((p0 & 0xffffu) + add), t1 = lds64((p0 >> 16) + add);
        if (COEF == 2) { t0 = dM(val[pp], t0); t1 = dM(val[pp + 1], t1); }
        acc = dA(acc, t0); acc = dA(acc, t1);
        pp += 64;
    }
    if (W & 1) {                                     // the odd last entry of a slice is stored alone, one u16 per lane
        const int pos = pp - lane;
        double t = lds64(ld_idx<GIDX>(ia, gidx, pos) + add);
        if (COEF == 2) t = dM(val[pos], t);
        acc = dA(acc, t);
    }
    return acc;
}
// two products over the same pattern (the two column products of the rhs, LP.cpp:875-878): operands at idx and idx + add2
template <int COEF, bool GIDX>
__device__ __forceinline__ void ell_dot2(uint32_t ia, const u16 *__restrict__ gidx, const double *__restrict__ val1,
                                         const double *__restrict__ val2, int base, int lane, int W, uint32_t add2, double &o1, double &o2) {
    double acc1 = 0.0, acc2 = 0.0;
    int k = 0, pp = base + 2 * lane;
    if (W >= 2) {
        uint32_t p0 = ld_pair<GIDX>(ia, gidx, pp);
#pragma unroll 1
        for (;;) {
            const uint32_t i0 = p0 & 0xffffu, i1 = p0 >> 16;
            double t0 = lds64(i0), t1 = lds64(i1), u0 = lds64(i0 + add2), u1 = lds64(i1 + add2);
            if (COEF == 2) { t0 = dM(val1[pp], t0); t1 = dM(val1[pp + 1], t1); u0 = dM(val2[pp], u0); u1 = dM(val2[pp + 1], u1); }
            k += 2; pp += 64;
            const bool more = (k + 2 <= W);
            if (more) p0 = ld_pair<GIDX>(ia, gidx, pp);
            acc1 = dA(acc1, t0); acc2 = dA(acc2, u0); acc1 = dA(acc1, t1); acc2 = dA(acc2, u1);
            if (!more) break;
        }
    }
    if (k < W) {
        const int pos = pp - lane;
        const uint32_t i0 = ld_idx<GIDX>(ia, gidx, pos);
        double t0 = lds64(i0), u0 = lds64(i0 + add2);
        if (COEF == 2) { t0 = dM(val1[pos], t0); u0 = dM(val2[pos], u0); }
        acc1 = dA(acc1, t0); acc2 = dA(acc2, u0);
    }
    o1 = acc1; o2 = acc2;
}

// Builds one orientation of the padded sliced-ELL image from the compressed one.  Block-cooperative (any blockDim, all
// threads must call).  ptr/idx(/val): compressed arrays (outer = row for the row image, column for the column image),
// perm: slot -> outer index, inv: inner index -> slot of that inner index (its position in the gathered vector),
// base: byte offset of the gathered vector, zoff: byte offset of the shared 0.0.  s_w: shared scratch of >= count/32 + 2 ints.
__device__ __forceinline__ void build_ell(const u16 *ptr, const u16 *idx, const double *val, const u16 *perm, int count,
                                          const u16 *inv, int base, int zoff, u16 *o_sptr, u16 *o_idx, double *o_val, int *s_w) {
    const int ns = (count + 31) >> 5;
    for (int w = threadIdx.x; w <= ns; w += blockDim.x) s_w[w] = 0;
    __syncthreads();
    for (int s = threadIdx.x; s < count; s += blockDim.x) {
        const int o = perm[s];
        atomicMax(&s_w[s >> 5], (int)ptr[o + 1] - (int)ptr[o]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int w = 0; w < ns; ++w) { int wd = s_w[w]; s_w[w] = run; o_sptr[w] = (u16)run; run += wd; }
        s_w[ns] = run; o_sptr[ns] = (u16)run;
    }
    __syncthreads();
    for (int s = threadIdx.x; s < ns * 32; s += blockDim.x) {
        const int w = s >> 5;
        const int W = s_w[w + 1] - s_w[w];
        const int slice0 = s_w[w] * 32, l = s & 31;
        int L = 0, beg = 0;
        if (s < count) { const int o = perm[s]; beg = ptr[o]; L = ptr[o + 1] - beg; }
        for (int k = 0; k < W; ++k) {
            const int pos = ell_pos(slice0, l, k, W);
            if (k < L) { o_idx[pos] = (u16)(base + 8 * (int)inv[idx[beg + k]]); if (o_val) o_val[pos] = val[beg + k]; }
            else { o_idx[pos] = (u16)zoff; if (o_val) o_val[pos] = 0.0; }
        }
    }
    __syncthreads();
}
// builds both orientations; rinv / cinv: shared scratch (u16) of >= m / n entries
__device__ __forceinline__ void build_ell_image(const BatchView &bv, int inst, const InstState *st, int n, int m, const u16 *rowptr,
                                                const u16 *colidx, const u16 *colptr, const u16 *rowidx, bool unit, long long ov,
                                                u16 *rinv, u16 *cinv, int *s_w) {
    const EllLayout EL = ell_layout(st->n0, st->m0, st->rcap, st->ccap);
    unsigned char *ell = bv.pat + bv.off_pat[inst];
    const u16 *rperm = reinterpret_cast<const u16 *>(ell + EL.o_rperm), *cperm = reinterpret_cast<const u16 *>(ell + EL.o_cperm);
    for (int s = threadIdx.x; s < m; s += blockDim.x) rinv[rperm[s]] = (u16)s;
    for (int s = threadIdx.x; s < n; s += blockDim.x) cinv[cperm[s]] = (u16)s;
    __syncthreads();
    build_ell(rowptr, colidx, unit ? nullptr : bv.val_r + ov, rperm, m, cinv, bv.sbase, bv.sbase + zero_off(bv.cap),
              reinterpret_cast<u16 *>(ell + EL.o_rsptr), reinterpret_cast<u16 *>(ell + EL.o_ridx),
              unit ? nullptr : bv.ev_r + bv.off_evr[inst], s_w);
    build_ell(colptr, rowidx, unit ? nullptr : bv.val_c + ov, cperm, n, rinv, bv.sbase + gather_base(bv.cap), bv.sbase + zero_off(bv.cap),
              reinterpret_cast<u16 *>(ell + EL.o_csptr), reinterpret_cast<u16 *>(ell + EL.o_cidx),
              unit ? nullptr : bv.ev_c + bv.off_evc[inst], s_w);
}

// shared-memory carve-up (all dynamic: the window kernel has no static shared memory, so its dynamic region starts at the
// same shared-window address as the probe kernel's -- BatchView::sbase -- which the ELL image offsets are relative to) -------
enum ScalarSlot : int { SB_RHO1 = 0, SB_RHO2, SB_RHO4, SB_PRHO1, SB_PRHO2, SB_PRHO4, SB_GAMMA, SB_RATIO, SB_STD, SB_CUR, SB_BEST,
                        SB_OBJLEN, SB_COUNT };
struct Smem {
    uint32_t G;      // [cap]   operand of E v in column-slot order (x or p); doubles as a third chain-major buffer; G + zero_off = 0.0
    uint32_t T1;     // [mp+2]  operand of E^T w in row-slot order: E v (scaled by rho4 in the unit case) / rho4 (f - y3)
    uint32_t R0, R1; // [4*CH]  chain-major reduction operands; R1 also holds the copy of z4 gathered by E^T z4
    uint32_t tab;    // [tab_len] unit case: 1 / precond_diag as a function of the column length
    uint32_t cst;    // [T] 8 bytes per thread: chain-major staging offsets (uint16, relative to R0) of the thread's columns
    double *sc;      // [16]    reduction results / broadcast scalars
    double *blk;     // [2][SB_COUNT] ADMM-level scalars (rho's, gamma, objective bookkeeping), double-buffered per iteration
    double *ring;    // [16]    tail of obj_list
    double *part;    // [2][4*NW] fast-mode partial sums
    int *ctl;        // [8]     work item, reduction-warp index, status of the iteration, PCG code, end of the slice
    double *ev_r, *ev_c, *r4v;    // ELL-order values: [evr_elems], [evc_elems], [evc_elems] (non-unit only)
    unsigned char *pat;
    uint64_t *bar;
};
__host__ __device__ inline size_t smem_bytes(int cap, int np, int mp, int pat_bytes, int evr_elems, int evc_elems, int nwarps, int tab_len) {
    size_t d = (size_t)((mp + 3) & ~1) + 8 * (size_t)chain_stride(np) + (size_t)((tab_len + 1) & ~1) + 32 * (size_t)nwarps + 16 +
               2 * SB_COUNT + 16 + 2 * 4 * (size_t)nwarps + 4 + (size_t)evr_elems + 2 * (size_t)evc_elems;
    return (size_t)gather_base(cap) + d * sizeof(double) + (size_t)pat_bytes + 16 + 256;
}
__device__ __forceinline__ Smem carve(unsigned char *base, int cap, int np, int mp, int pat_bytes, int evr_elems, int evc_elems, int nwarps,
                                      int tab_len) {
    Smem s;
    const uint32_t b32 = smem_u32(base);
    s.G = b32;
    double *d0 = reinterpret_cast<double *>(base + gather_base(cap));
    double *d = d0;
    const int CH = chain_stride(np);
    s.T1 = b32 + gather_base(cap); d += (mp + 3) & ~1;
    s.R0 = s.T1 + (uint32_t)(d - d0) * 8u; d += 4 * CH;
    s.R1 = s.T1 + (uint32_t)(d - d0) * 8u; d += 4 * CH;
    s.tab = s.T1 + (uint32_t)(d - d0) * 8u; d += (tab_len + 1) & ~1;
    s.cst = s.T1 + (uint32_t)(d - d0) * 8u; d += 32 * nwarps;
    s.sc = d; d += 16;
    s.blk = d; d += 2 * SB_COUNT;
    s.ring = d; d += 16;
    s.part = d; d += 2 * 4 * nwarps;
    s.ctl = reinterpret_cast<int *>(d); d += 4;
    s.ev_r = d; d += evr_elems;
    s.ev_c = d; d += evc_elems;
    s.r4v = d; d += evc_elems;
    s.pat = reinterpret_cast<unsigned char *>(d);
    // the image is followed by the mbarrier and 256 readable bytes: the unconditional offset prefetch of the SpMV loops may
    // read (never use) that far past the image's last array
    s.bar = reinterpret_cast<uint64_t *>(s.pat + pat_bytes);
    return s;
}
// reports the shared-window address at which the dynamic shared memory of a kernel WITHOUT static shared memory starts
static __global__ void lp_probe_kernel(int *out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    *out = (int)smem_u32(smem_raw);
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

// End-of-iteration bookkeeping of ADMM_lp_iters / _l2f (LP.cpp:931-1011), executed by ONE thread: stop test on x - y1 / x - y2,
// rho / gamma schedule, objective history + std stop, binary objective.  Reads the scalar block of this iteration, writes the
// block the next iteration (or the write-back) will read; returns the status.
static __device__ __noinline__ int admm_bookkeep(const Params &pr, const double *cur, double *nxt, double *ring, double sx, double sd1, double sd2,
                                          double sobj, double scur, bool guard, bool rho_step) {
    double rho1 = cur[SB_RHO1], rho2 = cur[SB_RHO2], rho4 = cur[SB_RHO4], prho1 = cur[SB_PRHO1], prho2 = cur[SB_PRHO2],
           prho4 = cur[SB_PRHO4], gamma = cur[SB_GAMMA], ratio = cur[SB_RATIO], std_obj = cur[SB_STD], cur_obj = cur[SB_CUR],
           best = cur[SB_BEST];
    long long obj_len = __double_as_longlong(cur[SB_OBJLEN]);
    int status = RUNNING;
    double temp0 = sqrt(sx);                                                 // :931
    if (!(temp0 > 2.2204e-16)) temp0 = 2.2204e-16;
    const double c1 = dD(sqrt(sd1), temp0), c2 = dD(sqrt(sd2), temp0);
    if (c1 <= pr.stop_threshold && c2 <= pr.stop_threshold && guard) status = STOP_Y;   // :934 / :1504 (break before anything else)
    else {
        if (rho_step) {                                                      // :951-970
            prho1 = rho1; prho2 = rho2; prho4 = rho4;
            rho1 = dM(pr.learning_fact, rho1); rho2 = dM(pr.learning_fact, rho2); rho4 = dM(pr.learning_fact, rho4);
            const double g = dM(gamma, pr.gamma_factor);
            gamma = (g < 1.0) ? 1.0 : g;
            ratio = dS(pr.learning_fact, 1.0);
        }
        double so = std_obj;                                                 // :972-976 (unchanged while size < history)
        if (obj_len + 1 >= (long long)pr.history_size) so = std_obj_after_push(ring, obj_len, sobj, pr.history_size);
        ring[obj_len & 15] = sobj;
        obj_len++;
        std_obj = so;
        if (std_obj <= pr.std_threshold) status = STOP_STD;                  // :977
        else {
            cur_obj = scur;                                                  // :1001-1005
            if (best >= cur_obj) best = cur_obj;                             // :1006-1009
        }
    }
    nxt[SB_RHO1] = rho1; nxt[SB_RHO2] = rho2; nxt[SB_RHO4] = rho4; nxt[SB_PRHO1] = prho1; nxt[SB_PRHO2] = prho2; nxt[SB_PRHO4] = prho4;
    nxt[SB_GAMMA] = gamma; nxt[SB_RATIO] = ratio; nxt[SB_STD] = std_obj; nxt[SB_CUR] = cur_obj; nxt[SB_BEST] = best;
    nxt[SB_OBJLEN] = __longlong_as_double(obj_len);
    return status;
}

// =====================================================================================================================
// The window kernel.  T threads, EPT slots (columns and rows) per thread: max(n, m) <= T*EPT.
// UNIT: all stored values of E are 1.0 (pattern-only matrices; rho4*E^T is one scalar; 1/precond_diag depends on the column
//       length only and is looked up in a small shared table).
// FAST: tree reductions + FMA (not bit-identical).
// =====================================================================================================================
#ifndef LPB_CTAS_128
#define LPB_CTAS_128 7
#endif
#ifndef LPB_CTAS_256
#define LPB_CTAS_256 4
#endif
#ifndef LPB_CTAS_512
#define LPB_CTAS_512 2
#endif
enum ParkSlot : int { PK_Y1 = 0, PK_Y2, PK_Z1, PK_Z2, PK_B, PK_X, PK_Y3, PK_Z4, PK_F, PK_COUNT };
constexpr int lpb_ctas(int T, int EPT) { return T <= 128 ? LPB_CTAS_128 : (T <= 256 ? (EPT <= 2 ? LPB_CTAS_256 : 3) : LPB_CTAS_512); }

template <int T, int EPT, bool UNIT, bool FAST>
__global__ void __launch_bounds__(T, lpb_ctas(T, EPT))
lp_admm_window_kernel(BatchView bv, Params pr, Launch la) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = T / 32, CAP = T * EPT;
    constexpr int CE = UNIT ? 0 : 2;     // coefficient mode of every sparse product
    const Smem S = carve(smem_raw, CAP, la.np, la.mp, la.pat_bytes, la.evr_elems, la.evc_elems, NW, la.tab_len);
    const int CH = chain_stride(la.np);
    const uint32_t z4_add = S.R1 - S.T1;                     // T1 operand address -> address of the z4 copy in R1
    const uint32_t g3_off = S.G - S.R0;                      // R0 staging address -> same chain-major position in the G region
    const uint32_t r1_off = S.R1 - S.R0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t tma_phase = 0;

    if (S.G != (uint32_t)bv.sbase) {                          // the ELL image holds absolute shared-window addresses
        if (tid == 0) atomicExch(la.error, 1);
        return;
    }
    if (tid == 0) {
        mbar_init(S.bar, 1); fence_mbar_init();
        S.ctl[1] = la.sm_rank ? (atomicAdd(la.sm_rank + smid(), 1) & (NW - 1)) : (NW - 1);
        S.ctl[5] = -2;                                                                  // sliced queue: nothing to hand back yet
        sts64(S.G + zero_off(CAP), 0.0); sts64(S.G + zero_off(CAP) + 8, 0.0);         // the shared padding operand
        sts64(S.R1 - 16, 0.0); sts64(S.R1 - 8, 0.0);                                  // padding operand of the z4 copy (zero + z4_add)
    }
    __syncthreads();
    // The warp that runs the sequential reductions (RW) is a different physical warp in CTAs that share an SM, so the
    // reduction warps do not pile up on one warp scheduler.  Slots are dealt by VIRTUAL warp: RW is always virtual warp
    // NW-1, whose slices hold the shortest rows / columns.
    const int rw = S.ctl[1];
    const bool is_rw = (warp == rw);
    const int vw = (warp + NW - 1 - rw) & (NW - 1);
    const int vt = vw * 32 + lane;
    const int rq = lane >> 2;                                 // reduction group of this lane (reduction warp only)
    const bool keeper = FAST ? (tid == 0) : (is_rw && lane == 0);   // the thread that does the scalar bookkeeping
    double *park = la.park + (size_t)blockIdx.x * PK_COUNT * CAP + vt;   // slot s of this thread: park[PK * CAP + e * T]
    int part_sel = 0;
    (void)part_sel; (void)rq; (void)g3_off; (void)r1_off;

    for (;;) {
        // Work items.  Plain mode: ticket w of the atomic counter is instance work[w].  SLICED mode (la.slice > 0, plain solves of a
        // whole batch): the queue is a ring of instance ids; a CTA runs at most `slice` iterations of the instance it popped and, if
        // the instance is still running, appends it to the ring again -- every instance advances at about the same pace, so when
        // the queue drains the CTAs finish within one slice of each other instead of one whole solve (the tail of a launch with
        // thousands of iterations per instance).  The window boundary carries the full solver state, so the iterates do not change.
        if (tid == 0) {
            if (la.slice > 0) {                                    // hand back / retire the instance this CTA has just left
                const int again = S.ctl[5];
                if (again >= 0) { const int slot = atomicAdd(la.tail, 1); if (slot < la.ring_cap) atomicExch(la.ring + slot, again); }
                else if (again == -1) atomicAdd(la.finished, 1);
                S.ctl[5] = -2;
            }
            const int t = atomicAdd(la.counter, 1);
            int item = -1;
            if (la.slice <= 0) { if (t < la.n_work) item = la.work ? la.work[t] : t; }
            else if (t < la.ring_cap) {
                volatile int *rg = la.ring;
                for (;;) {
                    const int v = rg[t];
                    if (v >= 0) { item = v; break; }
                    if (*(volatile int *)la.finished >= la.n_work) break;      // every instance is done: no more tickets will come
                    __nanosleep(256);
                }
                __threadfence();                                               // acquire: the producer's state writes are visible
            }
            S.ctl[0] = item;
        }
        __syncthreads();
        const int inst = S.ctl[0];
        __syncthreads();
        if (inst < 0) break;
        InstState *stp = bv.st + inst;
        if ((la.skip_done && __ldcg(&stp->done)) || stp->n == 0) {          // uniform per CTA
            if (la.slice > 0 && tid == 0) atomicAdd(la.finished, 1);
            continue;
        }

        // ---------------- stage the instance --------------------------------------------------------------------
        const int n = stp->n, m = stp->m;
        const EllLayout PL = ell_layout(stp->n0, stp->m0, stp->rcap, stp->ccap);
        // an image larger than the shared-memory budget of this launch keeps its last array (the column offsets) in global memory
        const bool spill = PL.o_rperm > la.pat_bytes;
        const unsigned char *gimg = bv.pat + bv.off_pat[inst];
        if (tid == 0) {
            fence_proxy_async();  // order earlier generic-proxy reads of the previous instance's blob before the overwrite
            mbar_expect_tx(S.bar, (uint32_t)(spill ? PL.o_cidx : PL.o_rperm));
            tma_load_1d(S.pat, gimg, (uint32_t)(spill ? PL.o_cidx : PL.o_rperm), S.bar);
        }
        const long long on = bv.off_n[inst], om = bv.off_m[inst];
        const u16 *g_rperm = reinterpret_cast<const u16 *>(gimg + PL.o_rperm);
        const u16 *g_cperm = reinterpret_cast<const u16 *>(gimg + PL.o_cperm);
        // registers across the whole window: x, chain-major staging address (in R0) of the owned columns, and 1/diag (unit
        // case: the column lengths, packed, that index the shared 1/diag table)
        double x[EPT], invd[UNIT ? 1 : EPT];
        uint32_t lens = 0;
        uint32_t cpk[2] = {0u, 0u};                          // chain-major staging offsets (bytes from R0) of the owned columns, 16 bits each
        LPB_FOR_E {
            const int s = vt + e * T;
            const bool in = s < n;
            const int j = in ? (int)g_cperm[s] : 0;
            cpk[e >> 1] |= ((uint32_t)((j & 3) * CH + (j >> 2)) * 8u) << (16 * (e & 1));
            // (mutable solver state is read past the L1: with the sliced queue another SM may have written it since this SM last saw it)
            x[e] = in ? __ldcg(bv.x + on + j) : 0.0;
            const double pd = in ? __ldcg(bv.Pd + on + j) : 1.0;
            const double iv = (pd != 0.0) ? dD(1.0, pd) : 1.0;   // value in use when rhoUpdated == 0 (Eigen: zero diagonal -> 1)
            if (UNIT) {
                const uint32_t L = in ? (uint32_t)bv.Esq[on + j] : 0u;      // unit case: Esq_j = number of stored entries of column j
                lens |= L << (8 * e);
                if (in) sts64(S.tab + 8u * L, iv);                            // same value from every column of that length
            } else invd[UNIT ? 0 : e] = iv;
            if (in) {
                park[PK_Y1 * CAP + e * T] = __ldcg(bv.y1 + on + j); park[PK_Y2 * CAP + e * T] = __ldcg(bv.y2 + on + j);
                park[PK_Z1 * CAP + e * T] = __ldcg(bv.z1 + on + j); park[PK_Z2 * CAP + e * T] = __ldcg(bv.z2 + on + j);
                park[PK_B * CAP + e * T] = bv.b[on + j];
            }
            if (s < m) {
                const int i = g_rperm[s];
                park[PK_Y3 * CAP + e * T] = __ldcg(bv.y3 + om + i); park[PK_Z4 * CAP + e * T] = __ldcg(bv.z4 + om + i); park[PK_F * CAP + e * T] = bv.f[om + i];
            }
        }
        sts_u32x2(S.cst + 8u * (uint32_t)tid, make_uint2(cpk[0], cpk[1]));   // read back (by this thread only) at every staging step
        // staging addresses of this thread's columns in R0; kept in shared memory between uses (registers are scarce in PCG)
#define LPB_CST_LOAD()                                                                                                \
        uint32_t cst[EPT];                                                                                            \
        if (!FAST) { const uint2 cw_ = lds_u32x2(S.cst + 8u * (uint32_t)tid);                                         \
                     LPB_FOR_E cst[e] = S.R0 + (((((e) >> 1) ? cw_.y : cw_.x) >> (16 * ((e) & 1))) & 0xffffu); }      \
        else { LPB_FOR_E cst[e] = 0u; }
#define LPB_INVD(e) (UNIT ? lds64(S.tab + 8u * ((lens >> (8 * (e))) & 0xffu)) : invd[UNIT ? 0 : (e)])
        double D = __ldcg(&stp->D), r4s = __ldcg(&stp->r4s);
        int rhoUpdated = __ldcg(&stp->rhoUpdated);
        int cg_total = 0, admm_total = 0;
        int cur = 0;                                          // which scalar block the iteration about to start reads
        if (tid == 0) {
            double *b0 = S.blk;
            b0[SB_RHO1] = __ldcg(&stp->rho1); b0[SB_RHO2] = __ldcg(&stp->rho2); b0[SB_RHO4] = __ldcg(&stp->rho4); b0[SB_PRHO1] = __ldcg(&stp->prho1);
            b0[SB_PRHO2] = __ldcg(&stp->prho2); b0[SB_PRHO4] = __ldcg(&stp->prho4); b0[SB_GAMMA] = __ldcg(&stp->gamma); b0[SB_RATIO] = __ldcg(&stp->ratio);
            b0[SB_STD] = __ldcg(&stp->std_obj); b0[SB_CUR] = __ldcg(&stp->cur_obj); b0[SB_BEST] = __ldcg(&stp->best_bin_obj);
            b0[SB_OBJLEN] = __longlong_as_double(__ldcg(&stp->obj_len));
        }
        if (tid < 16) S.ring[tid] = __ldcg(&stp->obj_ring[tid]);
        const int evr_used = UNIT ? 0 : 32 * stp->rcap, evc_used = UNIT ? 0 : 32 * stp->ccap;
        if (!UNIT) {
            const long long er = bv.off_evr[inst], ec = bv.off_evc[inst];
            for (int k = tid; k < evr_used; k += T) S.ev_r[k] = bv.ev_r[er + k];
            for (int k = tid; k < evc_used; k += T) { S.ev_c[k] = bv.ev_c[ec + k]; S.r4v[k] = __ldcg(bv.r4v + ec + k); }
        }
        mbar_wait(S.bar, tma_phase); tma_phase ^= 1;
        const uint32_t pat_a = smem_u32(S.pat);
        const uint32_t rsptr = pat_a + PL.o_rsptr, csptr = pat_a + PL.o_csptr, ridx = pat_a + PL.o_ridx, cidx = pat_a + PL.o_cidx;
        const u16 *gcidx = reinterpret_cast<const u16 *>(gimg + PL.o_cidx);   // used instead of cidx when the image is spilled
        const double pow_n = bv.pow_tab[n];               // std::pow(n, 1.0/p), p = 2 (LP.cpp:427)
        const int nsr = (m + 31) >> 5, nsc = (n + 31) >> 5;

        // E v on the row slots of this thread -> T1 (slot order).  SCALE (unit case): T1 = rho4E^T-scalar * (E v), the operand
        // the column product needs (LP.cpp:115-162: rho4_E_transpose * (E * v) multiplies every stored value r4s with (E v)_i).
#define LPB_ROW_SPMV(SCALE)                                                                                           \
    do {                                                                                                              \
        _Pragma("unroll 1") for (int e = 0; e < EPT; ++e) {                                                           \
            const int sl = vw + e * NW;                                                                               \
            if (sl < nsr) {                                                                                           \
                const int w0 = lds16(rsptr + 2 * sl), W = LPB_UNIFORM((int)lds16(rsptr + 2 * sl + 2) - w0);           \
                double acc = ell_dot<CE, false>(ridx, nullptr, S.ev_r, w0 * 32, lane, W, 0u);                        \
                if (SCALE) acc = dM(r4s, acc);                                                                        \
                if (vt + e * T < m) sts64(S.T1 + 8u * (uint32_t)(vt + e * T), acc);                                   \
            }                                                                                                         \
        }                                                                                                             \
    } while (0)
        // column product of slot e of this thread from T1 (COEF 0: T1 already carries the rho4 factor; 2: rho4E^T values)
#define LPB_COL_DOT(e, OUT)                                                                                           \
    do {                                                                                                              \
        const int sl = vw + (e) * NW;                                                                                 \
        OUT = 0.0;                                                                                                    \
        if (sl < nsc) {                                                                                               \
            const int w0 = lds16(csptr + 2 * sl), W = LPB_UNIFORM((int)lds16(csptr + 2 * sl + 2) - w0);               \
            OUT = spill ? ell_dot<CE, true>(0u, gcidx, S.r4v, w0 * 32, lane, W, 0u)                                  \
                        : ell_dot<CE, false>(cidx, nullptr, S.r4v, w0 * 32, lane, W, 0u);                            \
        }                                                                                                             \
    } while (0)
        // fast-mode CTA-wide sums (alternating partial buffers: one barrier per call)
#define LPB_BLOCK_SUM(K, V)                                                                                           \
    do { block_sum<K, NW>(V, S.part + part_sel * 4 * NW); part_sel ^= 1; } while (0)

        int status = RUNNING;
        int iter = la.slice > 0 ? __ldcg(&stp->iter) : la.iter_start;          // sliced: continue where the previous slice of this instance stopped
        if (tid == 0) S.ctl[4] = la.slice > 0 ? min(iter + la.slice, la.iter_end) : la.iter_end;   // end of this slice (read from shared memory: registers are scarce)
#define it_end (*(volatile int *)(S.ctl + 4))
        int cc = 0;
        const bool lp_plain = (!la.l2f) && pr.guard_first_iter;
        double ynorm2 = 0.0;                                  // ||x + z2/rho2 - 1/2||^2 of the iteration about to start

        // ---- window prologue: E x -> T1 and ||y||^2 for the first iteration (later iterations get both from the tail
        //      of the previous one: same operands, same order -> same bits) ---------------------------------------------
        __syncthreads();                                      // image, ring, scalar block visible; previous instance's smem reads done
        if (iter < it_end) {
            double ysq[1] = {0.0};
            const double rho2 = S.blk[SB_RHO2];
            LPB_CST_LOAD();
            LPB_FOR_E {
                const int s = vt + e * T;
                if (s < n) {
                    const double y = dS(dA(x[e], dD(park[PK_Z2 * CAP + e * T], rho2)), 0.5);   // :815, :424
                    sts64(S.G + 8u * (uint32_t)s, x[e]);
                    if (FAST) ysq[0] = fma(y, y, ysq[0]);
                    else sts64(cst[e], dM(y, y));
                }
            }
            __syncthreads();
            LPB_ROW_SPMV(false);
            if (FAST) { LPB_BLOCK_SUM(1, ysq); ynorm2 = ysq[0]; }
            else if (is_rw) { const double v = warp_redux_cm1(S.R0, n, CH, 1); if (lane == 0) S.sc[5] = v; }
            __syncthreads();
            if (!FAST) ynorm2 = S.sc[5];
        }

        for (; iter < it_end; ++iter) {
            const double *blk = S.blk + cur * SB_COUNT;      // read-only during the iteration
            // ---- operator patch after a rho step (:851-866) first: the rhs uses the patched rho4 E^T --------------------
            if (iter != 0 && rhoUpdated) {
                const double ratio = blk[SB_RATIO];
                const double c12 = dM(ratio, dA(blk[SB_PRHO1], blk[SB_PRHO2]));
                const double c4 = dM(ratio, blk[SB_PRHO4]);
                D = dA(D, c12);
                LPB_FOR_E {
                    const int s = vt + e * T;
                    if (s < n) {
                        const int j = g_cperm[s];
                        double pd = __ldcg(bv.Pd + on + j);
                        pd = dA(pd, c12);
                        pd = dA(pd, dM(c4, bv.Esq[on + j]));
                        bv.Pd[on + j] = pd;
                    }
                }
                if (UNIT) r4s = dM(pr.learning_fact, r4s);
                else for (int k = tid; k < evc_used; k += T) S.r4v[k] = dM(pr.learning_fact, S.r4v[k]);
            }
            if (rhoUpdated) {                                                // preconditioner refresh (:883-890)
                LPB_FOR_E {
                    const int s = vt + e * T;
                    const double pd = (s < n) ? __ldcg(bv.Pd + on + g_cperm[s]) : 1.0;
                    const double iv = (pd != 0.0) ? dD(1.0, pd) : 1.0;
                    if (UNIT) { if (s < n) sts64(S.tab + 8u * ((lens >> (8 * e)) & 0xffu), iv); }
                    else invd[UNIT ? 0 : e] = iv;
                }
                rhoUpdated = 0;
            }
            // ---- y3 (:826-827) on the row slots; operands of the two column products of the rhs (:875, :878) ------------
            {
                const double rho4 = blk[SB_RHO4];
                LPB_FOR_E {
                    const int s = vt + e * T;
                    if (s < m) {
                        const double fi = park[PK_F * CAP + e * T], z4 = park[PK_Z4 * CAP + e * T];
                        const double t = dS(dS(fi, lds64(S.T1 + 8u * (uint32_t)s)), dD(z4, rho4));
                        const double y3 = (t < 0.0) ? 0.0 : t;
                        park[PK_Y3 * CAP + e * T] = y3;
                        const double d = dS(fi, y3);
                        sts64(S.T1 + 8u * (uint32_t)s, UNIT ? dM(r4s, d) : d);   // operand of rho4E^T (f - y3)
                        sts64(S.R1 + 8u * (uint32_t)s, z4);                      // operand of E^T z4
                    }
                }
            }
            __syncthreads();
            // ---- y1 (:806-809), y2 (:815-818, :424-427), rhs (:872-878); PCG warm start x = y1 (:892) ----------------
            // x[] is the PCG iterate from here on; the previous x (needed again only if PCG bails out in an early-fix window)
            // is parked.
            double rhs[EPT], r[EPT], p[EPT];
            double rn[1] = {0.0};
            {
                const double rho1 = blk[SB_RHO1], rho2 = blk[SB_RHO2];
                const double nrm = sqrt(ynorm2);
                const double den = dM(2.0, nrm);
                LPB_FOR_E {
                    const int s = vt + e * T;
                    double a = 0.0, c = 0.0;
                    const int sl = vw + e * NW;
                    if (sl < nsc) {
                        const int w0 = lds16(csptr + 2 * sl), W = LPB_UNIFORM((int)lds16(csptr + 2 * sl + 2) - w0);
                        if (spill) ell_dot2<CE, true>(0u, gcidx, S.r4v, S.ev_c, w0 * 32, lane, W, z4_add, a, c);
                        else ell_dot2<CE, false>(cidx, nullptr, S.r4v, S.ev_c, w0 * 32, lane, W, z4_add, a, c);
                    }
                    if (s < n) {
                        const double z1 = park[PK_Z1 * CAP + e * T], z2 = park[PK_Z2 * CAP + e * T], bj = park[PK_B * CAP + e * T];
                        const double t1 = dA(x[e], dD(z1, rho1));
                        const double y1 = (t1 > 1.0) ? 1.0 : ((t1 < 0.0) ? 0.0 : t1);
                        const double yp = dS(dA(x[e], dD(z2, rho2)), 0.5);
                        const double y2 = dA(dD(dM(yp, pow_n), den), 0.5);   // :427
                        park[PK_Y1 * CAP + e * T] = y1; park[PK_Y2 * CAP + e * T] = y2; park[PK_X * CAP + e * T] = x[e];
                        double t = dS(dA(dM(rho1, y1), dM(rho2, y2)), dA(dA(bj, z1), z2));
                        t = dA(t, a);
                        rhs[e] = dS(t, c);
                        x[e] = y1;
                    } else { rhs[e] = 0.0; x[e] = 0.0; }
                }
            }
            __syncthreads();                                                 // all reads of T1 / R1 (z4 copy) done
            { LPB_CST_LOAD();
            LPB_FOR_E {
                const int s = vt + e * T;
                if (s < n) {
                    sts64(S.G + 8u * (uint32_t)s, x[e]);
                    if (FAST) rn[0] = fma(rhs[e], rhs[e], rn[0]);
                    else sts64(cst[e], dM(rhs[e], rhs[e]));
                }
            } }
            __syncthreads();
            LPB_ROW_SPMV(UNIT);                                              // E x0
            // PCG scalars (parity mode): the keeper -- lane 0 of the reduction warp -- forms threshold, alpha, beta and the stop
            // decisions right after each reduction and publishes them; nobody else carries them in registers.
            //   sc[0] alpha | sc[2] beta | sc[8] threshold | sc[9] absNew | ctl[3] 0 = iterate, 1 = rhs is zero, 2 = converged
            if (!FAST && is_rw) {
                const double v = warp_redux_cm1(S.R0, n, CH, 1);                // rhs.squaredNorm() :277
                if (lane == 0) {
                    double threshold = dM(dM(pr.pcg_tol, pr.pcg_tol), v);        // :287
                    if (!(threshold > DBL_MIN)) threshold = DBL_MIN;
                    S.sc[8] = threshold;
                    S.ctl[3] = (v == 0.0) ? 1 : 0;                               // :279-284
                }
            }
            if (FAST) LPB_BLOCK_SUM(1, rn); else __syncthreads();
            double rr[2] = {0.0, 0.0};
            { LPB_CST_LOAD();
            LPB_FOR_E {
                const int s = vt + e * T;
                double acc;
                LPB_COL_DOT(e, acc);
                if (s < n) {
                    const double mv = dA(dA(0.0, dM(D, x[e])), acc);         // D v (+) R4ET (E v)   :115-162
                    r[e] = dS(rhs[e], mv);                                   // :273
                    p[e] = dM(LPB_INVD(e), r[e]);                            // :297
                    sts64(S.G + 8u * (uint32_t)s, p[e]);                     // (G was last read before the previous barrier)
                    if (FAST) { rr[0] = fma(r[e], r[e], rr[0]); rr[1] = fma(r[e], p[e], rr[1]); }
                    else { sts64(cst[e], dM(r[e], r[e])); sts64(cst[e] + r1_off, dM(r[e], p[e])); }
                } else { r[e] = 0.0; p[e] = 0.0; }
            } }
            int cg_code;
            double threshold = 0.0, absNew = 0.0;                            // fast mode only (parity: shared scalars)
            if (FAST) {
                LPB_BLOCK_SUM(2, rr);
                threshold = dM(dM(pr.pcg_tol, pr.pcg_tol), rn[0]);
                if (!(threshold > DBL_MIN)) threshold = DBL_MIN;
                absNew = rr[1];
                cg_code = (rn[0] == 0.0) ? 1 : ((rr[0] < threshold) ? 2 : 0);
            } else {
                __syncthreads();
                if (is_rw) {
                    const double v = warp_redux_cm1(rq == 0 ? S.R0 : S.R1, n, CH, 2);    // r.r :288, r.p :300
                    const double v1 = __shfl_sync(0xffffffffu, v, 4);
                    if (lane == 0) {
                        S.sc[9] = v1;                                            // absNew :300
                        S.ctl[3] = S.ctl[3] ? 1 : ((v < S.sc[8]) ? 2 : 0);       // :290-295
                    }
                }
                __syncthreads();
                cg_code = S.ctl[3];
            }
            int cg_it = 0;
            bool cg_fail = false;
            if (cg_code == 1) {                                              // :279-284
                LPB_FOR_E x[e] = 0.0;
            } else if (cg_code == 0) {
                while (cg_it < pr.pcg_maxiters) {                            // G holds p here
                    LPB_ROW_SPMV(UNIT);
                    __syncthreads();
                    double tmp[EPT];
                    double pq[1] = {0.0};
                    { LPB_CST_LOAD();
                    LPB_FOR_E {
                        const int s = vt + e * T;
                        double acc;
                        LPB_COL_DOT(e, acc);
                        if (s < n) {
                            tmp[e] = dA(dA(0.0, dM(D, p[e])), acc);          // :304
                            if (FAST) pq[0] = fma(p[e], tmp[e], pq[0]);
                            else sts64(cst[e], dM(p[e], tmp[e]));
                        } else tmp[e] = 0.0;
                    } }
                    double alpha;
                    if (FAST) { LPB_BLOCK_SUM(1, pq); alpha = dD(absNew, pq[0]); }
                    else {
                        __syncthreads();
                        if (is_rw) {
                            const double v = warp_redux_cm1(S.R0, n, CH, 1);    // p.dot(tmp) :306
                            if (lane == 0) S.sc[0] = dD(S.sc[9], v);            // alpha = absNew / p.tmp
                        }
                        __syncthreads();
                        alpha = S.sc[0];
                    }
                    if (pr.alpha_bailout && alpha < 0.0) { cg_fail = true; break; }  // :307
                    double rz[2] = {0.0, 0.0};
                    { LPB_CST_LOAD();
                    LPB_FOR_E {
                        const int s = vt + e * T;
                        if (FAST) {
                            x[e] = fma(alpha, p[e], x[e]);
                            r[e] = fma(-alpha, tmp[e], r[e]);
                        } else {
                            x[e] = dA(x[e], dM(alpha, p[e]));                 // :308
                            r[e] = dS(r[e], dM(alpha, tmp[e]));               // :310
                        }
                        if (s < n) {
                            const double zz = dM(LPB_INVD(e), r[e]);          // :320 (formed again after the reduction: same product)
                            if (FAST) { rz[0] = fma(r[e], r[e], rz[0]); rz[1] = fma(r[e], zz, rz[1]); }
                            else { sts64(cst[e], dM(r[e], r[e])); sts64(cst[e] + r1_off, dM(r[e], zz)); }
                        }
                    } }
                    double beta;
                    bool cg_done;
                    if (FAST) {
                        LPB_BLOCK_SUM(2, rz);
                        cg_done = rz[0] < threshold;
                        beta = dD(rz[1], absNew);
                        absNew = rz[1];
                    } else {
                        __syncthreads();
                        if (is_rw) {
                            const double v = warp_redux_cm1(rq == 0 ? S.R0 : S.R1, n, CH, 2);   // r.r :311, r.z :323
                            const double v1 = __shfl_sync(0xffffffffu, v, 4);
                            if (lane == 0) {
                                const int done = (v < S.sc[8]) ? 1 : 0;          // :315-318
                                S.ctl[3] = done;
                                if (!done) { S.sc[2] = dD(v1, S.sc[9]); S.sc[9] = v1; }   // beta = absNew / absOld :324
                            }
                        }
                        __syncthreads();
                        cg_done = S.ctl[3] != 0;
                        beta = S.sc[2];
                    }
                    if (cg_done) { cg_it++; break; }                         // :315-318
                    LPB_FOR_E {
                        const int s = vt + e * T;
                        if (s < n) {
                            const double zz = dM(LPB_INVD(e), r[e]);          // :320
                            p[e] = FAST ? fma(beta, p[e], zz) : dA(zz, dM(beta, p[e]));   // :325
                            sts64(S.G + 8u * (uint32_t)s, p[e]);
                        }
                    }
                    cg_it++;
                    __syncthreads();
                }
            }
            cg_total += cg_it;
            if (cg_fail && la.l2f) {                                         // :1450-1454 (x_sol keeps its old value)
                LPB_FOR_E { if (vt + e * T < n) x[e] = park[PK_X * CAP + e * T]; }
                status = STOP_CG;
                break;
            }
            // (plain loop: the return value is ignored, x_sol holds the partially updated iterate, :894)
            admm_total += 1;
            // ---- iterate history (:1472-1475) ----------------------------------------------------------------
            if ((la.l2f || la.record) && bv.hist_cap > 0) {
                if (cc < bv.hist_cap) {
                    double *h = bv.hist + bv.off_hist[inst] + (long long)cc * stp->n0;
                    LPB_FOR_E { const int s = vt + e * T; if (s < n) h[g_cperm[s]] = x[e]; }
                }
                cc++;
            }
            // ---- duals (:917-924); operands of (x-y1)^2, (x-y2)^2 (:932-933); x -> G for E x -------------------------
            double dd[2] = {0.0, 0.0};
            {
                const double gamma = blk[SB_GAMMA];
                const double g1 = dM(gamma, blk[SB_RHO1]), g2 = dM(gamma, blk[SB_RHO2]);
                LPB_CST_LOAD();
                LPB_FOR_E {
                    const int s = vt + e * T;
                    if (s < n) {
                        const double d1 = dS(x[e], park[PK_Y1 * CAP + e * T]), d2 = dS(x[e], park[PK_Y2 * CAP + e * T]);
                        park[PK_Z1 * CAP + e * T] = dA(park[PK_Z1 * CAP + e * T], dM(g1, d1));
                        park[PK_Z2 * CAP + e * T] = dA(park[PK_Z2 * CAP + e * T], dM(g2, d2));
                        sts64(S.G + 8u * (uint32_t)s, x[e]);
                        if (FAST) { dd[0] = fma(d1, d1, dd[0]); dd[1] = fma(d2, d2, dd[1]); }
                        else { sts64(cst[e], dM(d1, d1)); sts64(cst[e] + r1_off, dM(d2, d2)); }
                    }
                }
            }
            __syncthreads();
            LPB_ROW_SPMV(false);                                             // E x (raw): z4 update now, y3 of the next iteration
            if (!FAST && is_rw) {
                const double v = warp_redux_cm1(rq == 0 ? S.R0 : S.R1, n, CH, 2);
                if (lane == 0) S.sc[1] = v;
                if (lane == 4) S.sc[2] = v;
            }
            if (FAST) LPB_BLOCK_SUM(2, dd); else __syncthreads();
            // ---- z4 (:920-924); operands of x.x (:931), b.x (:972), b.1[x>=0.5] (:1001-1005) and of the NEXT iteration's
            //      ||y||^2 (:425; uses the duals just updated and the rho the next iteration will see) ----------------------
            const bool rho_step = ((iter + 1) % pr.rho_change_step == 0);
            const bool guard = lp_plain ? (iter != la.iter_start) : true;
            double ss[4] = {0.0, 0.0, 0.0, 0.0};
            {
                const double rho2 = blk[SB_RHO2];
                const double rho2n = rho_step ? dM(pr.learning_fact, rho2) : rho2;
                const double g4 = dM(blk[SB_GAMMA], blk[SB_RHO4]);
                const bool assign = lp_plain && (iter == la.iter_start);    // :920-921
                LPB_FOR_E {
                    const int s = vt + e * T;
                    if (s < m) {
                        const double t = dM(g4, dS(dA(lds64(S.T1 + 8u * (uint32_t)s), park[PK_Y3 * CAP + e * T]), park[PK_F * CAP + e * T]));
                        park[PK_Z4 * CAP + e * T] = assign ? t : dA(park[PK_Z4 * CAP + e * T], t);
                    }
                }
                LPB_CST_LOAD();
                LPB_FOR_E {
                    const int s = vt + e * T;
                    if (s < n) {
                        const double y = dS(dA(x[e], dD(park[PK_Z2 * CAP + e * T], rho2n)), 0.5);
                        const double bj = park[PK_B * CAP + e * T];
                        if (FAST) {
                            ss[0] = fma(x[e], x[e], ss[0]); ss[1] = fma(bj, x[e], ss[1]);
                            ss[2] += (x[e] >= 0.5) ? bj : 0.0; ss[3] = fma(y, y, ss[3]);
                        } else {
                            sts64(cst[e], x[e]); sts64(cst[e] + r1_off, bj); sts64(cst[e] + g3_off, dM(y, y));
                        }
                    }
                }
            }
            double *nxt = S.blk + (cur ^ 1) * SB_COUNT;
            if (FAST) {
                LPB_BLOCK_SUM(4, ss);
                ynorm2 = ss[3];
                if (keeper) S.ctl[2] = admm_bookkeep(pr, blk, nxt, S.ring, ss[0], dd[0], dd[1], ss[1], ss[2], guard, rho_step);
                __syncthreads();
            } else {
                __syncthreads();
                if (is_rw) {
                    // x.x | b.x | b.1[x>=0.5] | next ||y||^2 side by side
                    const uint32_t pa = (rq == 0) ? S.R0 : (rq == 3) ? S.G : S.R1;
                    const int mode = (rq == 2) ? 1 : (rq == 3) ? 2 : 0;
                    const double v = warp_redux_cm2(pa, S.R0, mode, n, CH, 4);
                    if ((lane & 3) == 0 && lane < 16) S.sc[rq == 0 ? 0 : (rq + 2)] = v;   // sc[0] x.x, sc[3] obj, sc[4] cur, sc[5] ||y||^2
                    __syncwarp();
                    if (keeper) S.ctl[2] = admm_bookkeep(pr, blk, nxt, S.ring, S.sc[0], S.sc[1], S.sc[2], S.sc[3], S.sc[4], guard, rho_step);
                }
                __syncthreads();
                ynorm2 = S.sc[5];
            }
            cur ^= 1;                                                        // the block just written is the current one now
            status = S.ctl[2];
            if (status == STOP_Y) break;                                     // :934 / :1504 (before the rho schedule)
            if (rho_step) rhoUpdated = 1;                                    // :951-970
            if (status == STOP_STD) break;                                   // :977
            // S.sc / S.ctl[2] are not written again before at least two more barriers
        }

        // ---------------- write the instance back ---------------------------------------------------------------
        __syncthreads();
        LPB_FOR_E {
            const int s = vt + e * T;
            if (s < n) {
                const int j = g_cperm[s];
                bv.x[on + j] = x[e]; bv.y1[on + j] = park[PK_Y1 * CAP + e * T]; bv.y2[on + j] = park[PK_Y2 * CAP + e * T];
                bv.z1[on + j] = park[PK_Z1 * CAP + e * T]; bv.z2[on + j] = park[PK_Z2 * CAP + e * T];
            }
            if (s < m) { const int i = g_rperm[s]; bv.y3[om + i] = park[PK_Y3 * CAP + e * T]; bv.z4[om + i] = park[PK_Z4 * CAP + e * T]; }
        }
        if (!UNIT) {
            const long long ec = bv.off_evc[inst];
            for (int k = tid; k < evc_used; k += T) bv.r4v[ec + k] = S.r4v[k];
        }
        if (tid < 16) stp->obj_ring[tid] = S.ring[tid];
        if (tid == 0) {
            const double *fb = S.blk + cur * SB_COUNT;
            stp->rho1 = fb[SB_RHO1]; stp->rho2 = fb[SB_RHO2]; stp->rho4 = fb[SB_RHO4]; stp->prho1 = fb[SB_PRHO1]; stp->prho2 = fb[SB_PRHO2];
            stp->prho4 = fb[SB_PRHO4]; stp->gamma = fb[SB_GAMMA]; stp->ratio = fb[SB_RATIO]; stp->std_obj = fb[SB_STD];
            stp->cur_obj = fb[SB_CUR]; stp->best_bin_obj = fb[SB_BEST]; stp->obj_len = __double_as_longlong(fb[SB_OBJLEN]);
            stp->D = D; stp->r4s = r4s; stp->rhoUpdated = rhoUpdated;
            stp->cg_iters = __ldcg(&stp->cg_iters) + cg_total; stp->admm_iters = __ldcg(&stp->admm_iters) + admm_total;
            stp->iter = iter; stp->status = status;
            int ret;
            if (la.l2f) ret = (status != RUNNING || __ldcg(&stp->norm_small)) ? 1 : 0;    // :1505, :1542, :1452, :1223
            else ret = (status == STOP_STD) ? 1 : 0;                             // :978
            stp->last_ret = ret;
            // the window driver stops calling once a call returned 1 (LP.trainer:521); the plain driver calls once
            stp->done = la.l2f ? ret : (status != RUNNING);
            if (la.l2f || la.record) { stp->xit_cols = cc; if (!la.l2f) stp->xit_rows = n; }
            S.ctl[5] = (status == RUNNING && iter < la.iter_end) ? inst : -1;   // sliced queue: ticket to append once the state is out
        }
        if (la.slice > 0) __threadfence();                                     // every thread's state writes precede the ticket
        __syncthreads();
        // (the ticket is appended at the top of the loop, before the next pop)
    }
#undef it_end
#undef LPB_ROW_SPMV
#undef LPB_COL_DOT
#undef LPB_BLOCK_SUM
#undef LPB_INVD
}

// =====================================================================================================================
// Set-up kernel: ADMM_lp_iters_init (LP.cpp:489-763) and/or update_expression (LP.cpp:2289-2404) for every instance.
// One CTA per instance, operating in HBM.  mode bit 0: initialise the iterate state; bit 1: rebuild the operator;
// bit 2: (re)build the padded sliced-ELL image.
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
    __shared__ u16 s_rinv[2048], s_cinv[2048];
    if (mode & 4) build_ell_image(bv, inst, st, n, m, rowptr, colidx, colptr, rowidx, unit, ov, s_rinv, s_cinv, s_w);
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
