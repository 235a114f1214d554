// Early-fixing policy network as bf16 tensor-core kernels for sm_100a.
//
// GraphAttentionEncoder / MLPEncoder forward (LP.mha:202-304) in eval mode for R variables x T tokens:
//   embed (CUDA cores, K = 10) -> per layer { QKV GEMM -> attention (T <= 32, 8 heads, d_k = 16; CUDA cores)
//   -> out-proj GEMM (+skip, BatchNorm folded) -> FF1 GEMM (+bias, ReLU) -> FF2 GEMM (+bias, +skip, BatchNorm folded) }
//   -> fc1 GEMM (K = T*128) -> fc2 GEMM -> head (fc3, fc4, sigmoid; CUDA cores).
// 97 % of the MACs are in the GEMMs, which run on the 5th-generation tensor cores: tcgen05.mma (kind::f16, bf16 x bf16
// -> fp32 in TMEM) issued by one elected thread, operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) through
// a 4-stage mbarrier ring, accumulators read back with tcgen05.ld and a fused epilogue
//   out = [ (acc + bias[n]) (+ residual[m][n]) ] -> [ReLU] -> [* scale[n] + shift[n]] -> bf16.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/lpbox_b200.h"

void lpbox_set_error(const std::string &s);

namespace {

constexpr int BM = 128, BN = 128, BK = 64, STAGES = 2;   // K is 128..2560: 2 stages x 32 KB -> 3 CTAs per SM overlap load / MMA / epilogue
constexpr int GEMM_THREADS = 192;          // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2..5: epilogue
constexpr uint32_t TMEM_COLS = 128;        // 128 lanes x 128 fp32 columns = one 128 x 128 accumulator tile
constexpr size_t STAGE_BYTES = (size_t)(BM + BN) * BK * 2;
static_assert(STAGES * STAGE_BYTES >= (size_t)BM * BN * 4, "the epilogue stages a fp32 tile in the operand ring");
constexpr size_t GEMM_SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t s2u(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t *b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s2u(b)), "r"(c)); }
__device__ __forceinline__ void mb_expect(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s2u(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_wait(uint64_t *b, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "LW:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LD;\n\t"
        "bra LW;\n\t"
        "LD:\n\t}" ::"r"(s2u(b)), "r"(parity) : "memory");
}
// one lane of a converged warp (the branch around it must be warp-uniform so that descriptors stay in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s2u(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(s2u(bar)) : "memory");
}
// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
// start address >> 4 in [0,14), LBO (16-byte units) in [16,30) = 1, SBO in [32,46) = 8 rows * 128 B = 1024 B,
// version = 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    const uint32_t lo = ((smem_addr >> 4) & 0x3FFF) | (1u << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
    return d;
}
// instruction descriptor for kind::f16 (cute::UMMA::InstrDescriptor): D = fp32 (c_format 1, bits [4,6)), A = B = bf16
// (format 1, bits [7,10) and [10,13)), both K-major (bits 15, 16 = 0), N >> 3 in [17,23), M >> 4 in [24,29)
__device__ __forceinline__ uint32_t make_idesc() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s2u(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct Epi {
    const float *bias;            // [N] or NULL
    const __nv_bfloat16 *resid;   // [M][ldc] or NULL
    const float *scale, *shift;   // [N] or NULL (folded BatchNorm, applied after the residual add)
    int relu;
};

__global__ void __launch_bounds__(GEMM_THREADS, 3)
gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, __nv_bfloat16 *__restrict__ C, int M,
                  int N, int K, int ldc, Epi ep) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B needs 1024 B
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    uint64_t *empty = full + STAGES;
    uint64_t *acc_full = empty + STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full + 1);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform role index
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;   // N-tiles of one M-tile are adjacent in launch order: A is read from HBM once
    const int nk = K / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mb_init(&full[s], 1); mb_init(&empty[s], 1); }
        mb_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM allocation by one warp (it also owns the deallocation)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    // Producer and MMA warps run their loops with all 32 lanes (warp-uniform control flow keeps the descriptors in uniform
    // registers); one elected lane issues the TMA / tcgen05 instructions.
    if (warp == 0) {
        // ---- TMA producer ----
        for (int kb = 0; kb < nk; ++kb) {
            const int s = kb % STAGES;
            if (kb >= STAGES) mb_wait(&empty[s], ((kb / STAGES) - 1) & 1);
            unsigned char *a = smem + (size_t)s * STAGE_BYTES, *b = a + (size_t)BM * BK * 2;
            if (elect_one()) {
                mb_expect(&full[s], (uint32_t)STAGE_BYTES);
                tma_2d(a, &mapA, kb * BK, m0, &full[s]);
                tma_2d(b, &mapB, kb * BK, n0, &full[s]);
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer ----
        const uint32_t idesc = make_idesc();
        for (int kb = 0; kb < nk; ++kb) {
            const int s = kb % STAGES;
            mb_wait(&full[s], (kb / STAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_addr = s2u(smem + (size_t)s * STAGE_BYTES), b_addr = a_addr + BM * BK * 2;
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)      // UMMA_K = 16 for bf16: advance 32 bytes inside the swizzled row
                    umma_f16(tmem, make_desc(a_addr + k * 32), make_desc(b_addr + k * 32), idesc, (kb | k) ? 1u : 0u);
                umma_commit(&empty[s]);   // frees the smem stage once the MMAs that read it have completed (implies fence::before_thread_sync)
                if (kb == nk - 1) umma_commit(acc_full);
            }
        }
    } else if (warp >= 2) {
        // ---- epilogue: warp w may touch TMEM lanes [32 (w % 4), +32) ----
        // Phase 1: accumulator rows -> fp32 staging tile in the (now idle) operand ring, 16-byte chunks XOR-swizzled by row.
        // Phase 2: the same warp walks its 32 rows; one instruction covers one whole row, so bias / residual / scale / shift
        // are per-lane constants and every global access is a full 128-byte line.
        const int q = warp & 3, r = q * 32 + lane;
        const int n = n0 + 4 * lane;
        uint2 rv[2][8];                                                        // residual rows in groups of 8, one group in flight ahead
        auto load_resid = [&](uint2 (&dst)[8], int g) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int m = m0 + q * 32 + g * 8 + i;
                dst[i] = m < M ? *reinterpret_cast<const uint2 *>(ep.resid + (size_t)m * ldc + n) : make_uint2(0u, 0u);
            }
        };
        if (ep.resid) load_resid(rv[0], 0);
        mb_wait(acc_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float *stage = reinterpret_cast<float *>(smem);                       // [128][128] fp32 = 64 KB = STAGES * STAGE_BYTES
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld32_nowait(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<uint4 *>(stage + r * BN + ((((c0 >> 2) + j) ^ (r & 7)) << 2)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        __syncwarp();
        float4 bi = make_float4(0.f, 0.f, 0.f, 0.f), sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ep.bias) bi = *reinterpret_cast<const float4 *>(ep.bias + n);
        if (ep.scale) { sc = *reinterpret_cast<const float4 *>(ep.scale + n); sh = *reinterpret_cast<const float4 *>(ep.shift + n); }
        auto finish_rows = [&](const uint2 (&res)[8], int g) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = q * 32 + g * 8 + i, m = m0 + rr;
                if (m >= M) continue;
                float4 a = *reinterpret_cast<const float4 *>(stage + rr * BN + ((lane ^ (rr & 7)) << 2));
                a.x += bi.x; a.y += bi.y; a.z += bi.z; a.w += bi.w;
                if (ep.resid) {
                    const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&res[i].x)), f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&res[i].y));
                    a.x += f0.x; a.y += f0.y; a.z += f1.x; a.w += f1.y;
                }
                if (ep.relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
                a.x = a.x * sc.x + sh.x; a.y = a.y * sc.y + sh.y; a.z = a.z * sc.z + sh.z; a.w = a.w * sc.w + sh.w;
                uint2 o;
                *reinterpret_cast<__nv_bfloat162 *>(&o.x) = __floats2bfloat162_rn(a.x, a.y);
                *reinterpret_cast<__nv_bfloat162 *>(&o.y) = __floats2bfloat162_rn(a.z, a.w);
                *reinterpret_cast<uint2 *>(C + (size_t)m * ldc + n) = o;
            }
        };
        if (ep.resid) load_resid(rv[1], 1);
        finish_rows(rv[0], 0);
        if (ep.resid) load_resid(rv[0], 2);
        finish_rows(rv[1], 1);
        if (ep.resid) load_resid(rv[1], 3);
        finish_rows(rv[0], 2);
        finish_rows(rv[1], 3);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}

// ---- fused feed-forward sublayer -------------------------------------------------------------------------------------------
// out = BN( X + W2 relu(W1 X + b1) + b2 )  (LP.mha:140-160: SkipConnection(Linear 128->512, ReLU, Linear 512->128) + Normalization)
// in ONE persistent kernel: the [rows][512] hidden activation never leaves the SM.  A CTA owns a PAIR of 128-row tiles (so
// every weight chunk fetched from L2 feeds 256 rows) and walks the hidden dimension in 8 chunks of 64:
//     acc1[t][b] (TMEM, 64 cols)  = X[t] . W1[chunk]^T             tcgen05.mma M128 N64, K = 128
//     H[t][b]    (smem, 16 KB)    = bf16(relu(acc1[t][b] + b1))    chunk-epilogue warps: tcgen05.ld -> cvt.relu -> 128B-swizzled K-major smem
//     acc2[t]    (TMEM, 128 cols) += H[t][b] . W2[:, chunk]^T      tcgen05.mma M128 N128, K = 64
// with b = chunk & 1, so the tensor core runs GEMM-1 of chunk c+1 while the epilogue warps convert chunk c.  After the 8th chunk
// the drain warps read acc2[t], add the bias, apply the folded BatchNorm and store bf16.  The residual never touches a CUDA core:
// acc2[t] is initialised by the tensor core as X[t] . I (exact in fp32), so X is read from HBM exactly once.
// TMEM: 2 tiles x (2 x 64 + 128) = 512 columns.  smem: X[2] 64 KB + H[2][2] 64 KB + W1[2] 32 KB + W2[2] 32 KB.
// Warp roles: 0 = TMA producer, 1 = MMA issuer / TMEM owner, 2..9 = chunk epilogue (4 per tile), 10..17 = drain (4 per tile).
constexpr int FF_THREADS = 32 * 18;
constexpr int FF_NC = 8;                         // hidden chunks of 64
constexpr uint32_t FF_TILE = 128 * 128 * 2;      // X tile: [128][128] bf16 = two [128][64] swizzled boxes
constexpr uint32_t FF_BOX = 128 * 64 * 2;        // one [128][64] swizzled box (H chunk, W2 chunk); the W1 chunk is two [64][64] boxes
constexpr uint32_t FF_IDN = 64 * 64 * 2;          // 64 x 64 identity (bf16, swizzled): the residual is added by the tensor core, acc2 += X . I
constexpr size_t FF_SMEM = 2 * (size_t)FF_TILE + 8 * (size_t)FF_BOX + FF_IDN + 1024 /*align*/ + (512 + 3 * 128) * 4 + 256 /*barriers*/;

__device__ __forceinline__ uint32_t make_idesc_n(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mb_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s2u(bar)) : "memory"); }
// bf16x2 {lo = relu(a), hi = relu(b)} in one instruction
__device__ __forceinline__ uint32_t cvt_relu_bf16x2(float a, float b) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}

__global__ void __launch_bounds__(FF_THREADS, 1)
ff_fused_tcgen05(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapW2,
                 const __nv_bfloat16 *__restrict__ X, __nv_bfloat16 *__restrict__ out, int M, const float *__restrict__ b1,
                 const float *__restrict__ b2, const float *__restrict__ s2, const float *__restrict__ t2) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char *sX = smem;                             // [tile][FF_TILE]
    unsigned char *sH = smem + 2 * FF_TILE;               // [tile][buf][FF_BOX]
    unsigned char *sW1 = sH + 4 * FF_BOX;                 // [buf][FF_BOX]: two [64 rows][64 k] boxes
    unsigned char *sW2 = sW1 + 2 * FF_BOX;                // [buf][FF_BOX]: one [128 rows][64 k] box
    unsigned char *sI = sW2 + 2 * FF_BOX;                 // [64 n][64 k] identity, K-major, 128-byte swizzle
    float *sb1 = reinterpret_cast<float *>(sI + FF_IDN), *sb2 = sb1 + 512, *ss2 = sb2 + 128, *st2 = ss2 + 128;
    uint64_t *bars = reinterpret_cast<uint64_t *>(st2 + 128);
    uint64_t *x_full = bars, *x_empty = bars + 2, *w1_full = bars + 4, *w1_empty = bars + 6, *w2_full = bars + 8, *w2_empty = bars + 10,
             *acc1_full = bars + 12 /*[t*2+buf]*/, *h_full = bars + 16 /*[t*2+buf]*/, *acc2_full = bars + 20, *acc2_empty = bars + 22;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 24);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform role index
    const int n_pairs = (M + 255) / 256;

    for (int i = threadIdx.x; i < 512; i += FF_THREADS) sb1[i] = b1[i];
    for (int i = threadIdx.x; i < 128; i += FF_THREADS) { sb2[i] = b2[i]; ss2[i] = s2[i]; st2[i] = b2[i] * s2[i] + t2[i]; }   // (acc + b2) s2 + t2 = acc s2 + st2
    for (int i = threadIdx.x; i < (int)FF_IDN / 2; i += FF_THREADS) {     // element (n, k) lives in 16-byte chunk (k / 8) ^ (n & 7) of row n
        const int n = i >> 6, ch = (i >> 3) & 7, e = i & 7, k = ((ch ^ (n & 7)) << 3) + e;
        reinterpret_cast<unsigned short *>(sI)[i] = (k == n) ? (unsigned short)0x3F80 : (unsigned short)0;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        for (int t = 0; t < 2; ++t) {
            mb_init(&x_full[t], 1); mb_init(&x_empty[t], 1); mb_init(&w1_full[t], 1); mb_init(&w1_empty[t], 1); mb_init(&w2_full[t], 1);
            mb_init(&w2_empty[t], 1); mb_init(&acc2_full[t], 1); mb_init(&acc2_empty[t], 128);
        }
        for (int i = 0; i < 4; ++i) { mb_init(&acc1_full[i], 1); mb_init(&h_full[i], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const int my_pairs = blockIdx.x < n_pairs ? (n_pairs - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const uint32_t n_it = (uint32_t)my_pairs * FF_NC;           // chunk steps of this CTA

    // Producer and MMA warps run their loops with all 32 lanes (warp-uniform control flow keeps descriptors in uniform
    // registers); one elected lane issues the TMA / tcgen05 instructions.
    if (warp == 0) {
        {
            // ---- TMA producer: chunk step `it` uses weight buffers it & 1 (phase (it >> 1) & 1) ----
            for (uint32_t it = 0; it < n_it; ++it) {
                const int c = it % FF_NC, buf = it & 1;
                const uint32_t pc = it / FF_NC;
                const int pair = blockIdx.x + (int)pc * gridDim.x;
                if (c == 0) {
                    for (int t = 0; t < 2; ++t) {
                        if (pc > 0) mb_wait(&x_empty[t], (pc - 1) & 1);
                        if (elect_one()) {
                            mb_expect(&x_full[t], FF_TILE);
                            tma_2d(sX + t * FF_TILE, &mapX, 0, pair * 256 + t * 128, &x_full[t]);
                            tma_2d(sX + t * FF_TILE + FF_BOX, &mapX, 64, pair * 256 + t * 128, &x_full[t]);
                        }
                    }
                }
                if (it >= 2) mb_wait(&w1_empty[buf], ((it >> 1) - 1) & 1);
                if (elect_one()) {
                    mb_expect(&w1_full[buf], FF_BOX);
                    tma_2d(sW1 + buf * FF_BOX, &mapW1, 0, c * 64, &w1_full[buf]);
                    tma_2d(sW1 + buf * FF_BOX + FF_BOX / 2, &mapW1, 64, c * 64, &w1_full[buf]);
                }
                if (it >= 2) mb_wait(&w2_empty[buf], ((it >> 1) - 1) & 1);
                if (elect_one()) {
                    mb_expect(&w2_full[buf], FF_BOX);
                    tma_2d(sW2 + buf * FF_BOX, &mapW2, c * 64, 0, &w2_full[buf]);
                }
            }
        }
    } else if (warp == 1) {
        {
            // ---- MMA issuer: GEMM-1 of step `it` is issued BEFORE GEMM-2 of step it-1, so the tensor core works while the
            // epilogue warps convert step it-1 (at a pair boundary the order flips: the next X tile is still in flight) ----
            const uint32_t idesc1 = make_idesc_n(64), idesc2 = make_idesc_n(128);
            auto gemm1 = [&](uint32_t it) {
                const int c = it % FF_NC, buf = it & 1;
                const uint32_t pc = it / FF_NC;
                mb_wait(&w1_full[buf], (it >> 1) & 1);
                for (int t = 0; t < 2; ++t) {
                    if (c == 0) mb_wait(&x_full[t], pc & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    // acc1[t][buf] was drained by the chunk epilogue of step it-2 (h_full waited before GEMM-2 of that step)
                    const uint32_t a = s2u(sX + t * FF_TILE), b = s2u(sW1 + buf * FF_BOX), d = tmem + t * 128 + buf * 64;
                    if (elect_one()) {
#pragma unroll
                        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_f16(d, make_desc(a + kb * FF_BOX + k * 32), make_desc(b + kb * (FF_BOX / 2) + k * 32), idesc1, (kb | k) ? 1u : 0u);
                        umma_commit(&acc1_full[t * 2 + buf]);
                        if (c == FF_NC - 1) umma_commit(&x_empty[t]);
                        if (t == 1) umma_commit(&w1_empty[buf]);
                    }
                }
            };
            auto gemm2 = [&](uint32_t it) {
                const int c = it % FF_NC, buf = it & 1;
                const uint32_t pc = it / FF_NC;
                mb_wait(&w2_full[buf], (it >> 1) & 1);
                for (int t = 0; t < 2; ++t) {
                    mb_wait(&h_full[t * 2 + buf], (it >> 1) & 1);
                    if (c == 0 && pc > 0) mb_wait(&acc2_empty[t], (pc - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a = s2u(sH + (t * 2 + buf) * FF_BOX), b = s2u(sW2 + buf * FF_BOX), d = tmem + 256 + t * 128;
                    const uint32_t xa = s2u(sX + t * FF_TILE), ia = s2u(sI);
                    if (elect_one()) {
                        if (c == 0) {       // residual: acc2[t][:, kb*64 .. +64) = X[t][:, kb*64 .. +64) . I
#pragma unroll
                            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                                for (int k = 0; k < 4; ++k) umma_f16(d + kb * 64, make_desc(xa + kb * FF_BOX + k * 32), make_desc(ia + k * 32), idesc1, k ? 1u : 0u);
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_f16(d, make_desc(a + k * 32), make_desc(b + k * 32), idesc2, 1u);
                        if (c == FF_NC - 1) umma_commit(&acc2_full[t]);
                        if (t == 1) umma_commit(&w2_empty[buf]);
                    }
                }
            };
            for (uint32_t it = 0; it < n_it; ++it) {
                if (it % FF_NC == 0) { if (it > 0) gemm2(it - 1); gemm1(it); }
                else { gemm1(it); gemm2(it - 1); }
            }
            if (n_it > 0) gemm2(n_it - 1);
        }
    } else if (warp < 10) {
        // ---- chunk epilogue: acc1[t][buf] -> relu(+b1) -> bf16 -> H[t][buf] (K-major, 128-byte swizzle: 16-byte chunk j of row r at j ^ (r & 7)) ----
        const int t = (warp - 2) >> 2, q = warp & 3, row = q * 32 + lane;
        for (uint32_t it = 0; it < n_it; ++it) {
            const int c = it % FF_NC, buf = it & 1;
            // acc1_full also tells that GEMM-2 of step it-2 (the last reader of H[t][buf]) has completed: it was issued earlier
            mb_wait(&acc1_full[t * 2 + buf], (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t r[64];
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * 128 + buf * 64);
            tmem_ld32_nowait(taddr, r);
            tmem_ld32_nowait(taddr + 32, r + 32);
            tmem_ld_wait();
            unsigned char *hrow = sH + (t * 2 + buf) * FF_BOX + row * 128;
            const float4 *bb = reinterpret_cast<const float4 *>(sb1 + c * 64);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 c0 = bb[2 * j], c1 = bb[2 * j + 1];
                uint4 o;
                o.x = cvt_relu_bf16x2(__uint_as_float(r[8 * j]) + c0.x, __uint_as_float(r[8 * j + 1]) + c0.y);
                o.y = cvt_relu_bf16x2(__uint_as_float(r[8 * j + 2]) + c0.z, __uint_as_float(r[8 * j + 3]) + c0.w);
                o.z = cvt_relu_bf16x2(__uint_as_float(r[8 * j + 4]) + c1.x, __uint_as_float(r[8 * j + 5]) + c1.y);
                o.w = cvt_relu_bf16x2(__uint_as_float(r[8 * j + 6]) + c1.z, __uint_as_float(r[8 * j + 7]) + c1.w);
                *reinterpret_cast<uint4 *>(hrow + ((j ^ (row & 7)) << 4)) = o;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the tensor core
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mb_arrive(&h_full[t * 2 + buf]);
        }
    } else {
        // ---- drain: acc2[t] (= residual + FF) -> (+ b2, folded BatchNorm) -> bf16 -> global ----
        const int t = (warp - 10) >> 2, q = warp & 3, row = q * 32 + lane;
        for (int pc = 0; pc < my_pairs; ++pc) {
            const int pair = blockIdx.x + pc * gridDim.x;
            const long long m = (long long)pair * 256 + t * 128 + row;
            mb_wait(&acc2_full[t], pc & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int g = 0; g < 4; ++g) {
                uint32_t r[32];
                tmem_ld32_nowait(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(256 + t * 128 + g * 32), r);
                tmem_ld_wait();
                if (m < M) {
                    const float4 *sc4 = reinterpret_cast<const float4 *>(ss2 + g * 32), *sh4 = reinterpret_cast<const float4 *>(st2 + g * 32);
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) {
                        const float4 a0 = sc4[2 * j8], a1 = sc4[2 * j8 + 1], c0 = sh4[2 * j8], c1 = sh4[2 * j8 + 1];
                        uint4 o;
                        __nv_bfloat162 *op = reinterpret_cast<__nv_bfloat162 *>(&o);
                        op[0] = __floats2bfloat162_rn(__uint_as_float(r[j8 * 8 + 0]) * a0.x + c0.x, __uint_as_float(r[j8 * 8 + 1]) * a0.y + c0.y);
                        op[1] = __floats2bfloat162_rn(__uint_as_float(r[j8 * 8 + 2]) * a0.z + c0.z, __uint_as_float(r[j8 * 8 + 3]) * a0.w + c0.w);
                        op[2] = __floats2bfloat162_rn(__uint_as_float(r[j8 * 8 + 4]) * a1.x + c1.x, __uint_as_float(r[j8 * 8 + 5]) * a1.y + c1.y);
                        op[3] = __floats2bfloat162_rn(__uint_as_float(r[j8 * 8 + 6]) * a1.z + c1.z, __uint_as_float(r[j8 * 8 + 7]) * a1.w + c1.w);
                        *reinterpret_cast<uint4 *>(out + m * 128 + g * 32 + j8 * 8) = o;
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mb_arrive(&acc2_empty[t]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

// ---- small CUDA-core kernels ----------------------------------------------------------------------------------------------
// embedding: h0[m][j] = sum_f We[j][f] x[m][f] + c[t][j],  c[t] = We[:,5:10] pe[t] + b  (LP.mha:229-235)
// a thread owns 8 consecutive outputs (its 40 weights live in registers) and walks tokens with stride 16 * gridDim.x;
// the 16 threads of a token write one 256-byte row.
constexpr int EMB_TOK = 64;   // tokens per thread
__global__ void __launch_bounds__(256) embed_kernel(const float *__restrict__ x, long long Mtok, int T, const float *__restrict__ We5 /*[128][5]*/,
                                                    const float *__restrict__ cpos /*[T][128]*/, __nv_bfloat16 *__restrict__ h) {
    const int j0 = (threadIdx.x & 15) * 8, sub = threadIdx.x >> 4;          // 16 tokens in flight per block
    float w[8][5];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int f = 0; f < 5; ++f) w[j][f] = We5[(j0 + j) * 5 + f];
    const long long base = (long long)blockIdx.x * (16 * EMB_TOK);
#pragma unroll 2
    for (int i = 0; i < EMB_TOK; ++i) {
        const long long m = base + i * 16 + sub;
        if (m >= Mtok) break;
        const int t = (int)(m % T);
        const float *xm = x + m * 5;
        const float x0 = xm[0], x1 = xm[1], x2 = xm[2], x3 = xm[3], x4 = xm[4];
        const float4 c0 = *reinterpret_cast<const float4 *>(cpos + t * 128 + j0), c1 = *reinterpret_cast<const float4 *>(cpos + t * 128 + j0 + 4);
        float v[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += w[j][0] * x0 + w[j][1] * x1 + w[j][2] * x2 + w[j][3] * x3 + w[j][4] * x4;
        uint4 o;
        __nv_bfloat162 *op = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) op[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        *reinterpret_cast<uint4 *>(h + m * 128 + j0) = o;
    }
}

// attention for one variable per CTA: qkv [T][384] (q | k | v, head-major inside each third) -> heads [T][128]
// softmax(q k' / sqrt(16)) v per head (LP.mha:83-104).  q, k, v are converted to fp32 ONCE while staging (shared memory
// [T][388] floats, +4 padding against bank conflicts); thread t < 8*T handles (head, query) = (t / T, t % T) with
// 128-bit shared loads.
constexpr int ATT_LD = 388;
template <int TT>   // TT > 0: compile-time token count (score array stays in registers); TT == 0: runtime T <= 32
__global__ void __launch_bounds__(256) attention_kernel(const __nv_bfloat16 *__restrict__ qkv, int Trt, __nv_bfloat16 *__restrict__ heads) {
    const int T = TT > 0 ? TT : Trt;
    extern __shared__ __align__(16) float sf[];                                    // [T][ATT_LD]
    const long long var = blockIdx.x;
    const uint4 *src = reinterpret_cast<const uint4 *>(qkv + var * T * 384);      // 48 x 16 bytes per token
    for (int i = threadIdx.x; i < T * 48; i += blockDim.x) {
        const uint4 u = src[i];
        const __nv_bfloat162 *p2 = reinterpret_cast<const __nv_bfloat162 *>(&u);
        const int tok = i / 48, c = (i - tok * 48) * 8;
        float *d = sf + tok * ATT_LD + c;
        const float2 f0 = __bfloat1622float2(p2[0]), f1 = __bfloat1622float2(p2[1]), f2 = __bfloat1622float2(p2[2]), f3 = __bfloat1622float2(p2[3]);
        *reinterpret_cast<float4 *>(d) = make_float4(f0.x, f0.y, f1.x, f1.y);
        *reinterpret_cast<float4 *>(d + 4) = make_float4(f2.x, f2.y, f3.x, f3.y);
    }
    __syncthreads();
    if (threadIdx.x >= 8 * T) return;
    const int hd = threadIdx.x / T, i = threadIdx.x - hd * T;
    float4 q[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        q[e] = *reinterpret_cast<const float4 *>(sf + i * ATT_LD + hd * 16 + 4 * e);
        q[e].x *= 0.25f; q[e].y *= 0.25f; q[e].z *= 0.25f; q[e].w *= 0.25f;
    }
    float sc[TT > 0 ? TT : 32], mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < T; ++j) {
        const float4 *kp = reinterpret_cast<const float4 *>(sf + j * ATT_LD + 128 + hd * 16);
        float d = 0.0f;
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float4 k4 = kp[e]; d += q[e].x * k4.x + q[e].y * k4.y + q[e].z * k4.z + q[e].w * k4.w; }
        sc[j] = d;
        mx = fmaxf(mx, d);
    }
    float den = 0.0f;
#pragma unroll
    for (int j = 0; j < T; ++j) { sc[j] = __expf(sc[j] - mx); den += sc[j]; }
    const float inv = 1.0f / den;
    float4 o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < T; ++j) {
        const float4 *vp = reinterpret_cast<const float4 *>(sf + j * ATT_LD + 256 + hd * 16);
        const float p = sc[j] * inv;
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float4 v4 = vp[e]; o[e].x += p * v4.x; o[e].y += p * v4.y; o[e].z += p * v4.z; o[e].w += p * v4.w; }
    }
    uint4 w0, w1;
    __nv_bfloat162 *a0 = reinterpret_cast<__nv_bfloat162 *>(&w0), *a1 = reinterpret_cast<__nv_bfloat162 *>(&w1);
    a0[0] = __floats2bfloat162_rn(o[0].x, o[0].y); a0[1] = __floats2bfloat162_rn(o[0].z, o[0].w);
    a0[2] = __floats2bfloat162_rn(o[1].x, o[1].y); a0[3] = __floats2bfloat162_rn(o[1].z, o[1].w);
    a1[0] = __floats2bfloat162_rn(o[2].x, o[2].y); a1[1] = __floats2bfloat162_rn(o[2].z, o[2].w);
    a1[2] = __floats2bfloat162_rn(o[3].x, o[3].y); a1[3] = __floats2bfloat162_rn(o[3].z, o[3].w);
    uint4 *dst = reinterpret_cast<uint4 *>(heads + (var * T + i) * 128 + hd * 16);
    dst[0] = w0; dst[1] = w1;
}

// ---- attention on the tensor cores (warp-level mma.sync; the tiles are too small for tcgen05) ------------------------------------
// One CTA per variable, one warp per head: S = Q K^T (m16n8k16 bf16 -> fp32), softmax on the accumulator fragments, P (bf16) V.
// qkv of the variable is staged once in shared memory ([32 tokens][392] bf16, pad rows zeroed); fragments come from ldmatrix.
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t &r0, uint32_t &r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t addr, uint32_t &r0, uint32_t &r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"   // register-only: free to be scheduled
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&v);
}

constexpr int ATM_LD = 392;    // bf16 per staged row: 784 bytes = 49 x 16 -> ldmatrix rows land in distinct banks
template <int TT>
__global__ void __launch_bounds__(256) attention_mma_kernel(const __nv_bfloat16 *__restrict__ qkv, __nv_bfloat16 *__restrict__ heads) {
    constexpr int MT = (TT + 15) / 16, NT = (TT + 7) / 8, KS = (TT + 15) / 16, ROWS = 16 * MT;
    __shared__ __align__(16) __nv_bfloat16 sq[ROWS * ATM_LD];
    const long long var = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {   // stage [TT][384] (48 x 16-byte chunks per token), zero the pad rows
        const uint4 *src = reinterpret_cast<const uint4 *>(qkv + var * TT * 384);
        for (int i = threadIdx.x; i < ROWS * 48; i += 256) {
            const int tok = i / 48, c = i - tok * 48;
            *reinterpret_cast<uint4 *>(sq + tok * ATM_LD + c * 8) = tok < TT ? src[i] : make_uint4(0u, 0u, 0u, 0u);
        }
    }
    __syncthreads();
    const int hd = warp;                                   // 8 warps = 8 heads
    const uint32_t base = s2u(sq);
    // ---- S = Q K^T ----
    float S[MT][NT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        uint32_t a[4];
        ldsm_x4(base + (uint32_t)(((mt * 16 + (lane & 15)) * ATM_LD + hd * 16 + (lane >> 4) * 8) * 2), a[0], a[1], a[2], a[3]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            uint32_t b0, b1;
            ldsm_x2(base + (uint32_t)(((nt * 8 + (lane & 7)) * ATM_LD + 128 + hd * 16 + ((lane >> 3) & 1) * 8) * 2), b0, b1);
            S[mt][nt][0] = S[mt][nt][1] = S[mt][nt][2] = S[mt][nt][3] = 0.0f;
            mma_bf16_16816(S[mt][nt], a, b0, b1);
        }
    }
    // ---- softmax over the TT valid columns of each row (rows lane/4 and lane/4 + 8 of every m-tile) ----
    uint32_t P[MT][2 * KS][2];                              // bf16x2 A-fragments of P: [m-tile][n-tile (zero beyond NT)][row half]
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            float mx = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int col = nt * 8 + (lane & 3) * 2 + j;
                    float v = S[mt][nt][hf * 2 + j] * 0.25f;
                    v = col < TT ? v : -INFINITY;
                    S[mt][nt][hf * 2 + j] = v;
                    mx = fmaxf(mx, v);
                }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            float den = 0.0f;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int j = 0; j < 2; ++j) { const float e = __expf(S[mt][nt][hf * 2 + j] - mx); S[mt][nt][hf * 2 + j] = e; den += e; }
            den += __shfl_xor_sync(0xffffffffu, den, 1);
            den += __shfl_xor_sync(0xffffffffu, den, 2);
            const float inv = 1.0f / den;
#pragma unroll
            for (int nt = 0; nt < 2 * KS; ++nt)
                P[mt][nt][hf] = nt < NT ? pack_bf16x2(S[mt][nt][hf * 2] * inv, S[mt][nt][hf * 2 + 1] * inv) : 0u;
        }
    }
    // ---- O = P V ----
    float O[MT][2][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int dt = 0; dt < 2; ++dt) O[mt][dt][0] = O[mt][dt][1] = O[mt][dt][2] = O[mt][dt][3] = 0.0f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
        for (int dt = 0; dt < 2; ++dt) {
            uint32_t b0, b1;
            ldsm_x2_trans(base + (uint32_t)(((ks * 16 + (lane & 15)) * ATM_LD + 256 + hd * 16 + dt * 8) * 2), b0, b1);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const uint32_t a[4] = {P[mt][2 * ks][0], P[mt][2 * ks][1], P[mt][2 * ks + 1][0], P[mt][2 * ks + 1][1]};
                mma_bf16_16816(O[mt][dt], a, b0, b1);
            }
        }
    }
    // ---- O -> bf16 into this head's (dead) Q columns, then whole rows go out with 16-byte stores ----
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int dt = 0; dt < 2; ++dt)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int row = mt * 16 + hf * 8 + (lane >> 2);
                *reinterpret_cast<uint32_t *>(sq + row * ATM_LD + hd * 16 + dt * 8 + (lane & 3) * 2) = pack_bf16x2(O[mt][dt][hf * 2], O[mt][dt][hf * 2 + 1]);
            }
    __syncthreads();
    uint4 *dst = reinterpret_cast<uint4 *>(heads + var * TT * 128);
    for (int i = threadIdx.x; i < TT * 16; i += 256) {
        const int tok = i >> 4, c = i & 15;
        dst[i] = *reinterpret_cast<const uint4 *>(sq + tok * ATM_LD + c * 8);
    }
}

// ---- fused multi-head-attention sublayer ---------------------------------------------------------------------------------------
// out = BN( X + Wo . MHA(X) )  (LP.mha:58-122 + SkipConnection + Normalization) in ONE persistent kernel: the [rows][384] q|k|v
// activation and the [rows][128] head outputs never leave the SM.  A tile is VPT = 128 / T whole variables (T tokens each):
//     acc_qkv (TMEM, 384 cols) = X . Wqkv^T                      tcgen05.mma, 3 chunks of N = 128, K = 128
//     acc_o   (TMEM, 128 cols) = X . I                           the residual, added by the tensor core
//     sQ (smem)  = bf16(acc_qkv) of 4 heads at a time            worker warps: tcgen05.ld -> shared memory
//     sHd (smem) = softmax(q k'/4) v per (variable, head)        worker warps: mma.sync + ldmatrix, written as the swizzled A operand
//     acc_o += sHd . Wo^T                                        tcgen05.mma
//     out = bf16(acc_o * scale + shift)                          worker warps
// Warp roles: 0 = TMA producer, 1 = MMA issuer / TMEM owner, 2..13 = workers.  The QKV GEMM of tile t+1 is issued as soon as the
// workers have read acc_qkv of tile t, so it overlaps attention / out-projection / drain of tile t.
constexpr int MH_WORKERS = 12;           // worker warps: 3 per TMEM lane quarter (= per SM sub-partition)
constexpr int MH_THREADS = 32 * (2 + MH_WORKERS);
constexpr int MH_QLD = 200;              // bf16 per staged row (Q | K | V of 4 heads = 192, + 8: 400 bytes = 25 x 16 -> conflict-free ldmatrix)
constexpr int MH_QROWS = 136;            // 128 tile rows + the rows a padded m-tile of the last variable can touch
constexpr size_t MH_SMEM = 5 * (size_t)FF_TILE + FF_IDN + (size_t)MH_QROWS * MH_QLD * 2 + 2 * 128 * 4 + 256 + 1024;

// Two attention tasks (variable v = tsk >> 2 of the tile, head hh = tsk & 3 of the staged half) interleaved in one warp:
// S = q k' (mma.sync), softmax on the fragments, O = P v, O -> the swizzled A operand of the out-projection (box `hd_box`).
template <int TT>
__device__ __forceinline__ void mha_attention_pair(uint32_t sq_base, unsigned char *hd_box, int tsk0, int tsk1, bool valid1, int lane) {
    constexpr int MT = (TT + 15) / 16, NT = (TT + 7) / 8, KS = (TT + 15) / 16;
    const int r0[2] = {(tsk0 >> 2) * TT, (tsk1 >> 2) * TT}, hh[2] = {tsk0 & 3, tsk1 & 3};
    float S[2][MT][NT][4];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            uint32_t a[4];
            ldsm_x4(sq_base + (uint32_t)(((r0[u] + mt * 16 + (lane & 15)) * MH_QLD + hh[u] * 16 + (lane >> 4) * 8) * 2), a[0], a[1], a[2], a[3]);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                uint32_t b0, b1;
                ldsm_x2(sq_base + (uint32_t)(((r0[u] + nt * 8 + (lane & 7)) * MH_QLD + 64 + hh[u] * 16 + ((lane >> 3) & 1) * 8) * 2), b0, b1);
                S[u][mt][nt][0] = S[u][mt][nt][1] = S[u][mt][nt][2] = S[u][mt][nt][3] = 0.0f;
                mma_bf16_16816(S[u][mt][nt], a, b0, b1);
            }
        }
    uint32_t P[2][MT][2 * KS][2];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int hr = 0; hr < 2; ++hr) {
                if (mt * 16 + hr * 8 >= TT) {                 // these 8 rows are all padding: no softmax, P = 0
#pragma unroll
                    for (int nt = 0; nt < 2 * KS; ++nt) P[u][mt][nt][hr] = 0u;
                    continue;
                }
                float mx = -INFINITY;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int col = nt * 8 + (lane & 3) * 2 + j;
                        float x = S[u][mt][nt][hr * 2 + j] * 0.25f;
                        x = col < TT ? x : -INFINITY;
                        S[u][mt][nt][hr * 2 + j] = x;
                        mx = fmaxf(mx, x);
                    }
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                float den = 0.0f;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int j = 0; j < 2; ++j) { const float e = __expf(S[u][mt][nt][hr * 2 + j] - mx); S[u][mt][nt][hr * 2 + j] = e; den += e; }
                den += __shfl_xor_sync(0xffffffffu, den, 1);
                den += __shfl_xor_sync(0xffffffffu, den, 2);
                const float inv = 1.0f / den;
#pragma unroll
                for (int nt = 0; nt < 2 * KS; ++nt)
                    P[u][mt][nt][hr] = nt < NT ? pack_bf16x2(S[u][mt][nt][hr * 2] * inv, S[u][mt][nt][hr * 2 + 1] * inv) : 0u;
            }
    float O[2][MT][2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int dt = 0; dt < 2; ++dt) O[u][mt][dt][0] = O[u][mt][dt][1] = O[u][mt][dt][2] = O[u][mt][dt][3] = 0.0f;
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int dt = 0; dt < 2; ++dt) {
                uint32_t b0, b1;
                ldsm_x2_trans(sq_base + (uint32_t)(((r0[u] + ks * 16 + (lane & 15)) * MH_QLD + 128 + hh[u] * 16 + dt * 8) * 2), b0, b1);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const uint32_t a[4] = {P[u][mt][2 * ks][0], P[u][mt][2 * ks][1], P[u][mt][2 * ks + 1][0], P[u][mt][2 * ks + 1][1]};
                    mma_bf16_16816(O[u][mt][dt], a, b0, b1);
                }
            }
    // head output -> A operand of the out-projection (K-major, 128-byte swizzle; k = (4 hf + hh) * 16 + d lives in box hf)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        if (u == 1 && !valid1) break;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int hr = 0; hr < 2; ++hr) {
                const int ti = mt * 16 + hr * 8 + (lane >> 2);
                if (ti < TT) {
                    const int rr = r0[u] + ti;
#pragma unroll
                    for (int dt = 0; dt < 2; ++dt)
                        *reinterpret_cast<uint32_t *>(hd_box + rr * 128 + (((hh[u] * 2 + dt) ^ (rr & 7)) << 4) + (lane & 3) * 4) =
                            pack_bf16x2(O[u][mt][dt][hr * 2], O[u][mt][dt][hr * 2 + 1]);
                }
            }
    }
}

template <int TT>
__global__ void __launch_bounds__(MH_THREADS, 1)
mha_fused_tcgen05(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapWqkv, const __grid_constant__ CUtensorMap mapWo,
                  __nv_bfloat16 *__restrict__ out, int M, const float *__restrict__ scale, const float *__restrict__ shift) {
    constexpr int VPT = 128 / TT, TR = VPT * TT;
    constexpr int MT = (TT + 15) / 16, NT = (TT + 7) / 8, KS = (TT + 15) / 16;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char *sX = smem, *sW = sX + FF_TILE /*[2]*/, *sWo = sW + 2 * FF_TILE, *sHd = sWo + FF_TILE, *sI = sHd + FF_TILE;
    __nv_bfloat16 *sQ = reinterpret_cast<__nv_bfloat16 *>(sI + FF_IDN);
    float *ssc = reinterpret_cast<float *>(sQ + MH_QROWS * MH_QLD), *ssh = ssc + 128;
    uint64_t *bars = reinterpret_cast<uint64_t *>(ssh + 128);
    uint64_t *x_full = bars, *x_empty = bars + 1, *w_full = bars + 2 /*[2]*/, *w_empty = bars + 4 /*[2]*/, *wo_full = bars + 6, *qkv_full = bars + 7,
             *qkv_empty = bars + 8, *hd_full = bars + 9, *o_full = bars + 10, *acc_o_empty = bars + 11;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 12);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int n_tiles = (M + TR - 1) / TR;
    const int my_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    for (int i = threadIdx.x; i < 128; i += MH_THREADS) { ssc[i] = scale[i]; ssh[i] = shift[i]; }
    for (int i = threadIdx.x; i < (int)FF_IDN / 2; i += MH_THREADS) {
        const int n = i >> 6, ch = (i >> 3) & 7, e = i & 7, k = ((ch ^ (n & 7)) << 3) + e;
        reinterpret_cast<unsigned short *>(sI)[i] = (k == n) ? (unsigned short)0x3F80 : (unsigned short)0;
    }
    for (int i = threadIdx.x; i < MH_QROWS * MH_QLD / 2; i += MH_THREADS) reinterpret_cast<uint32_t *>(sQ)[i] = 0u;   // pad rows must stay finite
    for (int i = threadIdx.x; i < (int)FF_TILE / 4; i += MH_THREADS) reinterpret_cast<uint32_t *>(sHd)[i] = 0u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        mb_init(x_full, 1); mb_init(x_empty, 1); mb_init(&w_full[0], 1); mb_init(&w_full[1], 1); mb_init(&w_empty[0], 1); mb_init(&w_empty[1], 1);
        mb_init(wo_full, 1); mb_init(qkv_full, 1); mb_init(qkv_empty, 32 * MH_WORKERS); mb_init(hd_full, 32 * MH_WORKERS); mb_init(o_full, 1); mb_init(acc_o_empty, 32 * MH_WORKERS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        // ---- TMA producer ----
        if (my_tiles > 0 && elect_one()) {
            mb_expect(wo_full, FF_TILE);
            tma_2d(sWo, &mapWo, 0, 0, wo_full);
            tma_2d(sWo + FF_BOX, &mapWo, 64, 0, wo_full);
        }
        for (int i = 0; i < my_tiles; ++i) {
            const int row0 = (blockIdx.x + i * gridDim.x) * TR;
            if (i > 0) mb_wait(x_empty, (i - 1) & 1);
            if (elect_one()) {
                mb_expect(x_full, FF_TILE);
                tma_2d(sX, &mapX, 0, row0, x_full);
                tma_2d(sX + FF_BOX, &mapX, 64, row0, x_full);
            }
            for (int c = 0; c < 3; ++c) {
                const uint32_t cc = (uint32_t)i * 3 + c, buf = cc & 1;
                if (cc >= 2) mb_wait(&w_empty[buf], ((cc >> 1) - 1) & 1);
                if (elect_one()) {
                    mb_expect(&w_full[buf], FF_TILE);
                    tma_2d(sW + buf * FF_TILE, &mapWqkv, 0, c * 128, &w_full[buf]);
                    tma_2d(sW + buf * FF_TILE + FF_BOX, &mapWqkv, 64, c * 128, &w_full[buf]);
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer ----
        const uint32_t idesc128 = make_idesc_n(128), idesc64 = make_idesc_n(64);
        auto qkv_gemm = [&](int i) {
            mb_wait(x_full, i & 1);
            for (int c = 0; c < 3; ++c) {
                const uint32_t cc = (uint32_t)i * 3 + c, buf = cc & 1;
                mb_wait(&w_full[buf], (cc >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a = s2u(sX), b = s2u(sW + buf * FF_TILE), d = tmem + c * 128;
                if (elect_one()) {
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_f16(d, make_desc(a + kb * FF_BOX + k * 32), make_desc(b + kb * FF_BOX + k * 32), idesc128, (kb | k) ? 1u : 0u);
                    umma_commit(&w_empty[buf]);
                    if (c == 2) umma_commit(qkv_full);
                }
            }
        };
        auto residual = [&]() {       // acc_o[:, kb*64 .. +64) = X[:, kb*64 .. +64) . I ; afterwards the X buffer is free
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a = s2u(sX), b = s2u(sI), d = tmem + 384;
            if (elect_one()) {
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16(d + kb * 64, make_desc(a + kb * FF_BOX + k * 32), make_desc(b + k * 32), idesc64, k ? 1u : 0u);
                umma_commit(x_empty);
            }
        };
        if (my_tiles > 0) { qkv_gemm(0); residual(); }
        for (int i = 0; i < my_tiles; ++i) {
            if (i + 1 < my_tiles) { mb_wait(qkv_empty, i & 1); qkv_gemm(i + 1); }
            if (i == 0) mb_wait(wo_full, 0);
            mb_wait(hd_full, i & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {
                const uint32_t a = s2u(sHd), b = s2u(sWo), d = tmem + 384;
                if (elect_one()) {
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_f16(d, make_desc(a + kb * FF_BOX + k * 32), make_desc(b + kb * FF_BOX + k * 32), idesc128, 1u);
                    umma_commit(o_full);
                }
            }
            if (i + 1 < my_tiles) { mb_wait(acc_o_empty, i & 1); residual(); }
        }
    } else {
        // ---- workers ----
        const int w8 = warp - 2, q = warp & 3, wsub = w8 >> 2, row = q * 32 + lane;      // wsub = 0..2: the q / k / v part this warp stages
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        const uint32_t sq_base = s2u(sQ);
        // acc_qkv -> sQ: q | k | v of heads 4 hf .. 4 hf + 3; this warp copies the 64 columns of part `wsub` for its 32 rows
        auto stage = [&](int hf) {
#pragma unroll 1
            for (int g = 0; g < 2; ++g) {
                uint32_t r[32];
                tmem_ld32_nowait(lane_base + (uint32_t)(wsub * 128 + hf * 64 + g * 32), r);
                tmem_ld_wait();
                __nv_bfloat16 *dst = sQ + row * MH_QLD + wsub * 64 + g * 32;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4 *>(dst + 8 * j) =
                        make_uint4(pack_bf16x2(__uint_as_float(r[8 * j]), __uint_as_float(r[8 * j + 1])), pack_bf16x2(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3])),
                                   pack_bf16x2(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5])), pack_bf16x2(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7])));
            }
        };
        auto worker_bar = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(32 * MH_WORKERS) : "memory"); };
        if (my_tiles > 0) {
            mb_wait(qkv_full, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            stage(0);
        }
        for (int i = 0; i < my_tiles; ++i) {
            const long long row0 = (long long)(blockIdx.x + i * gridDim.x) * TR;
#pragma unroll 1
            for (int hf = 0; hf < 2; ++hf) {
                if (hf == 1) {
                    stage(1);
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");      // every read of acc_qkv is done: the next tile's
                    mb_arrive(qkv_empty);                                                 // QKV GEMM may overwrite it
                }
                worker_bar();
                // attention tasks (variable v of the tile, head 4 hf + hh), two at a time per warp for instruction-level parallelism
#pragma unroll 1
                for (int tsk = w8; tsk < VPT * 4; tsk += 2 * MH_WORKERS) {
                    const int tsk1 = tsk + MH_WORKERS;
                    mha_attention_pair<TT>(sq_base, sHd + hf * FF_BOX, tsk, tsk1 < VPT * 4 ? tsk1 : tsk, tsk1 < VPT * 4, lane);
                }
                worker_bar();                                                            // sQ is restaged next
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mb_arrive(hd_full);
            if (i + 1 < my_tiles) {                   // stage the first half of the next tile while the out-projection of this one runs
                mb_wait(qkv_full, (i + 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                stage(0);
            }
            // ---- drain: acc_o (= X + Wo heads) -> folded BatchNorm -> bf16 -> global; 8 groups of 16 columns over the 3 warps of a quarter ----
            mb_wait(o_full, i & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const long long m = row0 + row;
#pragma unroll 1
            for (int g = wsub * 3; g < (wsub == 2 ? 8 : wsub * 3 + 3); ++g) {
                uint32_t r[16];
                const int c0 = g * 16;
                tmem_ld16_nowait(lane_base + (uint32_t)(384 + c0), r);
                tmem_ld_wait();
                if (row < TR && m < M) {
                    const float4 *sc4 = reinterpret_cast<const float4 *>(ssc + c0), *sh4 = reinterpret_cast<const float4 *>(ssh + c0);
#pragma unroll
                    for (int j8 = 0; j8 < 2; ++j8) {
                        const float4 a0 = sc4[2 * j8], a1 = sc4[2 * j8 + 1], b0 = sh4[2 * j8], b1 = sh4[2 * j8 + 1];
                        uint4 o;
                        o.x = pack_bf16x2(__uint_as_float(r[j8 * 8 + 0]) * a0.x + b0.x, __uint_as_float(r[j8 * 8 + 1]) * a0.y + b0.y);
                        o.y = pack_bf16x2(__uint_as_float(r[j8 * 8 + 2]) * a0.z + b0.z, __uint_as_float(r[j8 * 8 + 3]) * a0.w + b0.w);
                        o.z = pack_bf16x2(__uint_as_float(r[j8 * 8 + 4]) * a1.x + b1.x, __uint_as_float(r[j8 * 8 + 5]) * a1.y + b1.y);
                        o.w = pack_bf16x2(__uint_as_float(r[j8 * 8 + 6]) * a1.z + b1.z, __uint_as_float(r[j8 * 8 + 7]) * a1.w + b1.w);
                        *reinterpret_cast<uint4 *>(out + m * 128 + c0 + j8 * 8) = o;
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mb_arrive(acc_o_empty);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

// head: a2 [R][128] (after fc2 + ReLU) -> fc3 (16) ReLU -> fc4 (1) -> sigmoid  (LP.mha:185-199)
__global__ void head_kernel(const __nv_bfloat16 *__restrict__ a2, long long R, const float *__restrict__ w3 /*[16][128]*/, const float *__restrict__ b3,
                            const float *__restrict__ w4 /*[16]*/, float b4, float *__restrict__ scores) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    float in[128];
    const __nv_bfloat16 *src = a2 + r * 128;
#pragma unroll
    for (int k = 0; k < 128; ++k) in[k] = __bfloat162float(src[k]);
    float logit = b4;
    for (int o = 0; o < 16; ++o) {
        float acc = b3[o];
#pragma unroll
        for (int k = 0; k < 128; ++k) acc += w3[o * 128 + k] * in[k];
        logit += w4[o] * fmaxf(acc, 0.0f);
    }
    scores[r] = 1.0f / (1.0f + __expf(-logit));
}

// ---- host side ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn get_encode() {
    static EncodeFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeFn)p;
    }
    return fn;
}
// row-major [rows][cols] bf16 matrix, box = [box_rows][64 cols], 128-byte swizzle, OOB -> zeros
bool make_map(CUtensorMap *map, const void *ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {cols, rows}, strides[1] = {cols * 2};
    cuuint32_t box[2] = {BK, box_rows}, estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int launch_gemm(cudaStream_t st, const __nv_bfloat16 *A, const __nv_bfloat16 *W, __nv_bfloat16 *C, long long M, int N, int K, const Epi &ep) {
    if (M <= 0) return 0;
    if (N % BN || K % BK) { lpbox_set_error("policy GEMM: N must be a multiple of 128 and K of 64"); return LPBOX_E_INVALID; }
    CUtensorMap ma, mb;
    if (!make_map(&ma, A, (uint64_t)M, (uint64_t)K, BM) || !make_map(&mb, W, (uint64_t)N, (uint64_t)K, BN)) { lpbox_set_error("cuTensorMapEncodeTiled failed"); return LPBOX_E_CUDA; }
    dim3 grid((unsigned)(N / BN), (unsigned)((M + BM - 1) / BM));
    gemm_bf16_tcgen05<<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(ma, mb, C, (int)M, N, K, N, ep);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { lpbox_set_error(std::string("policy GEMM launch: ") + cudaGetErrorString(e)); return LPBOX_E_CUDA; }
    return 0;
}


// per-device state: the opt-in shared-memory sizes are function attributes of the CURRENT device's context, and devices may
// differ in SM count -- both are set / read per device (lpbox_policy_create calls policy_prepare_device after cudaSetDevice)
static int g_sm_counts[64] = {0};
static bool g_attr_done[64] = {false};
static int sm_count_current() {
    int dev = 0; cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148;
    if (!g_sm_counts[dev]) { int c = 0; cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev); g_sm_counts[dev] = c > 0 ? c : 148; }
    return g_sm_counts[dev];
}
static bool policy_prepare_device() {
    int dev = 0; cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && g_attr_done[dev]) return true;
    bool ok = cudaFuncSetAttribute(gemm_bf16_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(ff_fused_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FF_SMEM) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(mha_fused_tcgen05<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MH_SMEM) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(mha_fused_tcgen05<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MH_SMEM) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(mha_fused_tcgen05<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MH_SMEM) == cudaSuccess;
    if (ok && dev >= 0 && dev < 64) g_attr_done[dev] = true;
    return ok;
}
int launch_ff_fused(cudaStream_t st, const __nv_bfloat16 *X, const __nv_bfloat16 *W1, const float *b1, const __nv_bfloat16 *W2, const float *b2,
                    const float *s2, const float *t2, __nv_bfloat16 *out, long long M) {
    if (M <= 0) return 0;
    if (!b1 || !b2 || !s2 || !t2) { lpbox_set_error("fused FF: bias / scale / shift vectors are required"); return LPBOX_E_INVALID; }
    CUtensorMap mx, m1, m2;
    if (!make_map(&mx, X, (uint64_t)M, 128, 128) || !make_map(&m1, W1, 512, 128, 64) || !make_map(&m2, W2, 128, 512, 128)) {
        lpbox_set_error("cuTensorMapEncodeTiled failed"); return LPBOX_E_CUDA;
    }
    const long long n_pairs = (M + 255) / 256;
    ff_fused_tcgen05<<<(unsigned)std::min<long long>(n_pairs, sm_count_current()), FF_THREADS, FF_SMEM, st>>>(mx, m1, m2, X, out, (int)M, b1, b2, s2, t2);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { lpbox_set_error(std::string("fused FF launch: ") + cudaGetErrorString(e)); return LPBOX_E_CUDA; }
    return 0;
}

template <int TT>
int launch_mha_fused_t(cudaStream_t st, const CUtensorMap &mx, const CUtensorMap &mq, const CUtensorMap &mo, __nv_bfloat16 *out, long long M,
                       const float *scale, const float *shift) {
    constexpr int TR = (128 / TT) * TT;
    const long long n_tiles = (M + TR - 1) / TR;
    mha_fused_tcgen05<TT><<<(unsigned)std::min<long long>(n_tiles, sm_count_current()), MH_THREADS, MH_SMEM, st>>>(mx, mq, mo, out, (int)M, scale, shift);
    return 0;
}
// out = (X + Wo . MHA(X)) * scale + shift for rows grouped in variables of T tokens (T = 20, 10 or 5)
int launch_mha_fused(cudaStream_t st, const __nv_bfloat16 *X, const __nv_bfloat16 *Wqkv, const __nv_bfloat16 *Wo, const float *scale,
                     const float *shift, __nv_bfloat16 *out, long long M, int T) {
    if (M <= 0) return 0;
    if (!scale || !shift || (T != 20 && T != 10 && T != 5) || M % T) { lpbox_set_error("fused MHA: T must be 20, 10 or 5 and rows a multiple of T"); return LPBOX_E_INVALID; }
    CUtensorMap mx, mq, mo;
    if (!make_map(&mx, X, (uint64_t)M, 128, 128) || !make_map(&mq, Wqkv, 384, 128, 128) || !make_map(&mo, Wo, 128, 128, 128)) {
        lpbox_set_error("cuTensorMapEncodeTiled failed"); return LPBOX_E_CUDA;
    }
    if (T == 20) launch_mha_fused_t<20>(st, mx, mq, mo, out, M, scale, shift);
    else if (T == 10) launch_mha_fused_t<10>(st, mx, mq, mo, out, M, scale, shift);
    else launch_mha_fused_t<5>(st, mx, mq, mo, out, M, scale, shift);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { lpbox_set_error(std::string("fused MHA launch: ") + cudaGetErrorString(e)); return LPBOX_E_CUDA; }
    return 0;
}

template <typename Tp>
Tp *dalloc(size_t n) { Tp *p = nullptr; return cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(Tp)) == cudaSuccess ? p : nullptr; }

}  // namespace

struct lpbox_policy {
    int device = 0, T = 0, L = 0;
    long long chunk = 0;     // variables per pass
    std::vector<void *> owned;
    float *We5 = nullptr, *cpos = nullptr, *w3 = nullptr, *b3 = nullptr, *w4 = nullptr;
    float b4 = 0;
    struct Layer { __nv_bfloat16 *Wqkv, *Wo, *W1, *W2; float *b1, *b2, *s1, *t1, *s2, *t2; };
    std::vector<Layer> layers;
    __nv_bfloat16 *Wfc1 = nullptr, *Wfc2 = nullptr;
    float *bfc1 = nullptr, *bfc2 = nullptr;
    __nv_bfloat16 *h = nullptr, *h2 = nullptr, *qkv = nullptr, *ff = nullptr, *a1 = nullptr, *a2 = nullptr;
    int64_t launches = 0;
    bool fused_mha = true;      // LPBOX_POLICY_UNFUSED_MHA=1 selects the three-kernel path (tests compare the two)
    bool failed = false;        // an allocation or upload failed during create
};

static __nv_bfloat16 *upload_bf16(lpbox_policy *p, const float *src, size_t n) {
    std::vector<__nv_bfloat16> tmp(n);
    for (size_t i = 0; i < n; ++i) tmp[i] = __float2bfloat16(src[i]);
    __nv_bfloat16 *d = dalloc<__nv_bfloat16>(n);
    if (d) { p->owned.push_back(d); if (cudaMemcpy(d, tmp.data(), n * 2, cudaMemcpyHostToDevice) != cudaSuccess) p->failed = true; }
    else p->failed = true;
    return d;
}
static float *upload_f32(lpbox_policy *p, const float *src, size_t n) {
    float *d = dalloc<float>(n);
    if (d) { p->owned.push_back(d); if (cudaMemcpy(d, src, n * 4, cudaMemcpyHostToDevice) != cudaSuccess) p->failed = true; }
    else p->failed = true;
    return d;
}

// Packed fp32 weights (see lpbox/policy_kernel.py: pack_policy): embed_w[128][10], embed_b[128], pe[T][5], per layer
// { Wqkv[384][128], Wo[128][128], bn1_scale[128], bn1_shift[128], W1[512][128], b1[512], W2[128][512], b2[128], bn2_scale[128],
// bn2_shift[128] }, fc1_w[256][T*128], fc1_b[256], fc2_w[128][256], fc2_b[128], fc3_w[16][128], fc3_b[16], fc4_w[16], fc4_b[1].
extern "C" lpbox_policy *lpbox_policy_create(int device, int tokens, int n_layers, const float *packed, int64_t n_packed, int64_t chunk_rows) {
    if (tokens <= 0 || tokens > 32 || n_layers < 0 || !packed) { lpbox_set_error("policy_create: bad arguments (tokens <= 32)"); return nullptr; }
    const int64_t per_layer = 384 * 128 + 128 * 128 + 256 + 512 * 128 + 512 + 128 * 512 + 128 + 256;
    const int64_t need = 128 * 10 + 128 + tokens * 5 + n_layers * per_layer + 256LL * tokens * 128 + 256 + 128 * 256 + 128 + 16 * 128 + 16 + 16 + 1;
    if (n_packed != need) { lpbox_set_error("policy_create: packed weight buffer has the wrong size"); return nullptr; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { lpbox_set_error("no CUDA device (there is no CPU fallback)"); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { lpbox_set_error("bad device"); return nullptr; }
    if (!get_encode()) { lpbox_set_error("cuTensorMapEncodeTiled not available"); return nullptr; }
    if (!policy_prepare_device()) { lpbox_set_error("policy_create: cannot set the shared-memory attributes of the policy kernels on this device"); return nullptr; }
    lpbox_policy *p = new lpbox_policy();
    p->device = device; p->T = tokens; p->L = n_layers; p->chunk = chunk_rows > 0 ? chunk_rows : 16384;
    p->fused_mha = getenv("LPBOX_POLICY_UNFUSED_MHA") == nullptr;
    const float *w = packed;
    const float *ew = w; w += 128 * 10;
    const float *eb = w; w += 128;
    const float *pe = w; w += tokens * 5;
    std::vector<float> we5(128 * 5), cpos((size_t)tokens * 128);
    for (int j = 0; j < 128; ++j) {
        for (int f = 0; f < 5; ++f) we5[j * 5 + f] = ew[j * 10 + f];
        for (int t = 0; t < tokens; ++t) { float a = eb[j]; for (int f = 0; f < 5; ++f) a += ew[j * 10 + 5 + f] * pe[t * 5 + f]; cpos[(size_t)t * 128 + j] = a; }
    }
    p->We5 = upload_f32(p, we5.data(), we5.size()); p->cpos = upload_f32(p, cpos.data(), cpos.size());
    for (int l = 0; l < n_layers; ++l) {
        lpbox_policy::Layer L;
        L.Wqkv = upload_bf16(p, w, 384 * 128); w += 384 * 128;
        L.Wo = upload_bf16(p, w, 128 * 128); w += 128 * 128;
        L.s1 = upload_f32(p, w, 128); w += 128; L.t1 = upload_f32(p, w, 128); w += 128;
        L.W1 = upload_bf16(p, w, 512 * 128); w += 512 * 128; L.b1 = upload_f32(p, w, 512); w += 512;
        L.W2 = upload_bf16(p, w, 128 * 512); w += 128 * 512; L.b2 = upload_f32(p, w, 128); w += 128;
        L.s2 = upload_f32(p, w, 128); w += 128; L.t2 = upload_f32(p, w, 128); w += 128;
        p->layers.push_back(L);
    }
    p->Wfc1 = upload_bf16(p, w, (size_t)256 * tokens * 128); w += (size_t)256 * tokens * 128; p->bfc1 = upload_f32(p, w, 256); w += 256;
    p->Wfc2 = upload_bf16(p, w, 128 * 256); w += 128 * 256; p->bfc2 = upload_f32(p, w, 128); w += 128;
    p->w3 = upload_f32(p, w, 16 * 128); w += 16 * 128; p->b3 = upload_f32(p, w, 16); w += 16;
    p->w4 = upload_f32(p, w, 16); w += 16; p->b4 = w[0];
    const size_t Mt = (size_t)p->chunk * tokens;
    p->h = dalloc<__nv_bfloat16>(Mt * 128); p->h2 = dalloc<__nv_bfloat16>(Mt * 128); p->qkv = dalloc<__nv_bfloat16>(Mt * 384);
    p->ff = dalloc<__nv_bfloat16>(Mt * 512); p->a1 = dalloc<__nv_bfloat16>((size_t)p->chunk * 256); p->a2 = dalloc<__nv_bfloat16>((size_t)p->chunk * 128);
    for (void *q : {(void *)p->h, (void *)p->h2, (void *)p->qkv, (void *)p->ff, (void *)p->a1, (void *)p->a2}) { if (q) p->owned.push_back(q); else p->failed = true; }
    if (p->failed) { lpbox_set_error("policy_create: out of device memory (or a weight upload failed)"); lpbox_policy_destroy(p); return nullptr; }
    return p;
}

extern "C" void lpbox_policy_destroy(lpbox_policy *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    cudaDeviceSynchronize();
    for (void *q : p->owned) cudaFree(q);
    delete p;
}

// scores_dev[r] = sigmoid(policy(input_dev[r])) for r < rows; input_dev: fp32 [rows][T*5] (the packed window history)
extern "C" int lpbox_policy_forward_dev(lpbox_policy *p, void *stream, const float *input_dev, int64_t rows, float *scores_dev) {
    if (!p || !input_dev || !scores_dev || rows < 0) return LPBOX_E_INVALID;
    if (cudaSetDevice(p->device) != cudaSuccess) return LPBOX_E_CUDA;
    if (!policy_prepare_device()) return LPBOX_E_CUDA;
    cudaStream_t st = (cudaStream_t)stream;
    const int T = p->T;
    for (int64_t r0 = 0; r0 < rows; r0 += p->chunk) {
        const long long R = std::min<long long>(p->chunk, rows - r0), Mt = R * T;
        embed_kernel<<<(unsigned)((Mt + 16 * EMB_TOK - 1) / (16 * EMB_TOK)), 256, 0, st>>>(input_dev + r0 * T * 5, Mt, T, p->We5, p->cpos, p->h);
        p->launches++;
        __nv_bfloat16 *h = p->h, *h2 = p->h2;
        for (auto &L : p->layers) {
            int rc;
            __nv_bfloat16 *hn = p->qkv;       // the sublayer output (h + MHA(h), BatchNorm folded) lives in the first Mt*128 elements of qkv
            if (p->fused_mha && (T == 20 || T == 10 || T == 5)) {
                rc = launch_mha_fused(st, h, L.Wqkv, L.Wo, L.s1, L.t1, hn, Mt, T); if (rc) return rc;
                p->launches -= 2;
            } else {
                rc = launch_gemm(st, h, L.Wqkv, p->qkv, Mt, 384, 128, Epi{nullptr, nullptr, nullptr, nullptr, 0}); if (rc) return rc;
                const unsigned at = ((8 * T + 31) / 32) * 32;
                const size_t asm_ = sizeof(float) * (size_t)T * ATT_LD;
                if (T == 20) attention_mma_kernel<20><<<(unsigned)R, 256, 0, st>>>(p->qkv, h2);
                else if (T == 10) attention_mma_kernel<10><<<(unsigned)R, 256, 0, st>>>(p->qkv, h2);
                else if (T == 5) attention_mma_kernel<5><<<(unsigned)R, 256, 0, st>>>(p->qkv, h2);
                else attention_kernel<0><<<(unsigned)R, at, asm_, st>>>(p->qkv, T, h2);   // CUDA-core version for other token counts
                rc = launch_gemm(st, h2, L.Wo, hn /*reuse as [Mt][128] scratch*/, Mt, 128, 128, Epi{nullptr, h, L.s1, L.t1, 0}); if (rc) return rc;
            }
            rc = launch_ff_fused(st, hn, L.W1, L.b1, L.W2, L.b2, L.s2, L.t2, h2, Mt); if (rc) return rc;
            std::swap(h, h2);
            p->launches += 4;
        }
        int rc = launch_gemm(st, h, p->Wfc1, p->a1, R, 256, T * 128, Epi{p->bfc1, nullptr, nullptr, nullptr, 1}); if (rc) return rc;
        rc = launch_gemm(st, p->a1, p->Wfc2, p->a2, R, 128, 256, Epi{p->bfc2, nullptr, nullptr, nullptr, 1}); if (rc) return rc;
        head_kernel<<<(unsigned)((R + 127) / 128), 128, 0, st>>>(p->a2, R, p->w3, p->b3, p->w4, p->b4, scores_dev + r0);
        p->launches += 3;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { lpbox_set_error(std::string("policy forward: ") + cudaGetErrorString(e)); return LPBOX_E_CUDA; }
    }
    return 0;
}

extern "C" int64_t lpbox_policy_launch_count(const lpbox_policy *p) { return p ? p->launches : -1; }

// plain GEMM entry for tests: C[M][N] (bf16) = A[M][K] (bf16) . W[N][K]^T (+ bias, ReLU)
extern "C" int lpbox_gemm_bf16_dev(void *stream, const void *A, const void *W, void *C, int64_t M, int N, int K, const float *bias, int relu) {
    if (!get_encode() || !policy_prepare_device()) { lpbox_set_error("cuTensorMapEncodeTiled / kernel attributes not available"); return LPBOX_E_CUDA; }
    return launch_gemm((cudaStream_t)stream, (const __nv_bfloat16 *)A, (const __nv_bfloat16 *)W, (__nv_bfloat16 *)C, M, N, K,
                       Epi{bias, nullptr, nullptr, nullptr, relu});
}

// the fused feed-forward sublayer alone (tests): out = ((X + W2 relu(W1 X + b1) + b2) * scale + shift), X/out bf16 [M][128],
// W1 bf16 [512][128], W2 bf16 [128][512]
extern "C" int lpbox_ff_fused_dev(void *stream, const void *X, const void *W1, const float *b1, const void *W2, const float *b2, const float *scale,
                                  const float *shift, void *out, int64_t M) {
    if (!get_encode() || !policy_prepare_device()) { lpbox_set_error("cuTensorMapEncodeTiled / kernel attributes not available"); return LPBOX_E_CUDA; }
    return launch_ff_fused((cudaStream_t)stream, (const __nv_bfloat16 *)X, (const __nv_bfloat16 *)W1, b1, (const __nv_bfloat16 *)W2, b2, scale, shift,
                           (__nv_bfloat16 *)out, M);
}

// the fused multi-head-attention sublayer alone (tests): out = (X + Wo MHA(X)) * scale + shift; X/out bf16 [M][128], M = variables * T,
// Wqkv bf16 [384][128] (q | k | v, head-major inside each third), Wo bf16 [128][128]; T = 20, 10 or 5
extern "C" int lpbox_mha_fused_dev(void *stream, const void *X, const void *Wqkv, const void *Wo, const float *scale, const float *shift, void *out,
                                   int64_t M, int T) {
    if (!get_encode() || !policy_prepare_device()) { lpbox_set_error("cuTensorMapEncodeTiled / kernel attributes not available"); return LPBOX_E_CUDA; }
    return launch_mha_fused((cudaStream_t)stream, (const __nv_bfloat16 *)X, (const __nv_bfloat16 *)Wqkv, (const __nv_bfloat16 *)Wo, scale, shift,
                            (__nv_bfloat16 *)out, M, T);
}
