// Fused backward of the single-head spatial attention (flash-style: no [S, S] matrix leaves the SM), tcgen05 + TMEM + TMA, sm_100a.
//
// What autograd derives for MultiScaleCrossSelfAttentionPRGCN.attention (/root/reference/models/layers.py:126-133) in loss.backward()
// (/root/reference/tools/run.py:78), for head dim 64 (level 1 of mscsa_prgcn.yaml, S = 4096: 94 % of the attention-backward work):
//   P = softmax_m(Q K^T)          dP = dO V^T          dS = P o (dP - rowdot),  rowdot[n] = <dO[n], P V[n]>
//   dV = P^T dO                   dK = dS^T Q          dQ = dS K
// Inputs are the forward's operands, the gradient of its output, the log-sum-exp rows the fused forward saved (hupr_attn_desc.lse)
// and rowdot (hupr_rowdot on the forward output); outputs are ADDED to fp32 buffers (dV of two attentions lands in one buffer).
//
// Formulation — "transposed scores", chosen so that every operand is consumed in the layout it already has:
//   * one CTA owns a block of 128 keys (K_j, V_j resident in shared memory, hi/lo) and walks the query blocks i of 128 rows;
//   * S^T = K_j Q_i^T and dP^T = V_j dO_i^T  (M = keys, N = queries, both operands K-major as loaded) -> TMEM, lanes = keys;
//   * the element-wise warps (thread = key row, half of the query columns) form P^T = exp(S^T - lse_i) and dS^T = P^T (dP^T - rowdot_i),
//     and write them back IN PLACE (per 32-query chunk) as packed hi/lo bf16 columns: they are the A operands, straight from tensor memory, of
//         dV_j += P^T dO_i     and     dK_j += dS^T Q_i        (B = dO_i / Q_i as MN-major operands: rows = queries = K axis)
//     — the A-from-TMEM path of the forward kernel and the MN-major descriptors of wgrad.cu;
//   * dS^T also goes to shared memory once, as a [keys][queries] tile that is the MN-major A operand (M = queries, K = keys) of
//         dQ_i = dS K_j        (B = K_j as an MN-major operand), accumulated fresh per query block and added to global memory
//     (every key block contributes to every dQ row: red.global.add.v4.f32);
//   * dV_j, dK_j accumulate in TMEM over all query blocks and are added to global memory at the end.
// TMEM columns: S^T/P^T [0,128), dP^T/dS^T [128,256), dV [256,320), dK [320,384), dQ [384,448).
// Shared memory: K, V 64 KB + Q, dO 64 KB + dS^T 64 KB + lse/rowdot 1 KB.
// Version 1 (attention_bwd_kernel, kept behind HUPR_ATTN_BWD_V1=1 for A/B) runs its stages back to back over 128-query blocks with one
// Q/dO buffer; version 2 (attention_bwd_kernel_v2, below, the default) pipelines 64-query sub-blocks.  Both are parity-tested against
// float64 autograd (tests/test_attention_bwd_gpu.py) and called by hupr_b200.training.TrainStep for the head-dim-64 level.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "hupr_internal.h"
#include "tc.cuh"
#include "split.cuh"

namespace hupr {

constexpr int AB_BK = 128;        // keys per CTA
constexpr int AB_BQ = 128;        // queries per iteration
constexpr int AB_D = 64;          // head dim
constexpr int AB_THREADS = 320;   // TMA warp, MMA warp, 8 element-wise warps
constexpr int AB_PLANE = 128 * 128;                 // bytes of one [128 rows][64 bf16] plane
constexpr int AB_SM_K = 0;                          // K_hi, K_lo
constexpr int AB_SM_V = 2 * AB_PLANE;               // V_hi, V_lo
constexpr int AB_SM_Q = 4 * AB_PLANE;               // Q_hi, Q_lo
constexpr int AB_SM_DO = 6 * AB_PLANE;              // dO_hi, dO_lo
constexpr int AB_SM_DS = 8 * AB_PLANE;              // dS^T hi: two 64-query atoms of [128 keys][128 B]; then lo
constexpr int AB_SM_VEC = 12 * AB_PLANE;            // float lse[128], rowdot[128]
constexpr int AB_SM_BAR = AB_SM_VEC + 1024;
constexpr int AB_SM_END = AB_SM_BAR + 128;
constexpr int AB_SMEM = AB_SM_END + 1024;           // + alignment slack
constexpr int AB_T_S = 0, AB_T_DP = 128, AB_T_DV = 256, AB_T_DK = 320, AB_T_DQ = 384;

struct AttnBwdParams {
    int nq;                       // query blocks
    int s;
    int q_off, k_off, v_off, do_off;
    const float* lse;
    const float* rowdot;
    float* dq; int dq_ld, dq_off;
    float* dk; int dk_ld, dk_off;
    float* dv; int dv_ld, dv_off;
};

__device__ __forceinline__ uint64_t ab_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {     // MN-major SWIZZLE_128B operand (see wgrad.cu)
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void ab_red_add_v4(float* dst, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(AB_THREADS, 1)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tmQ_hi, const __grid_constant__ CUtensorMap tmQ_lo,
                     const __grid_constant__ CUtensorMap tmK_hi, const __grid_constant__ CUtensorMap tmK_lo,
                     const __grid_constant__ CUtensorMap tmV_hi, const __grid_constant__ CUtensorMap tmV_lo,
                     const __grid_constant__ CUtensorMap tmO_hi, const __grid_constant__ CUtensorMap tmO_lo, const AttnBwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AB_SM_BAR);
    uint64_t* kv_full = bars;
    uint64_t* qd_full = bars + 1;      // Q_i, dO_i, lse_i, rowdot_i loaded
    uint64_t* s_full = bars + 2;       // S^T and dP^T in TMEM
    uint64_t* p_full = bars + 3;       // P^T, dS^T written (TMEM + shared memory): 256 arrivals
    uint64_t* dq_full = bars + 4;      // dV, dK, dQ MMAs of this query block retired: dQ may be read, Q/dO/dS buffers are free
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 5);
    float* s_lse = reinterpret_cast<float*>(smem + AB_SM_VEC);
    float* s_rd = s_lse + AB_BQ;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * AB_BK;
    const int b = blockIdx.y;

    if (threadIdx.x == 0) {
        mbar_init(kv_full, 1);
        mbar_init(qd_full, 1);
        mbar_init(s_full, 1);
        mbar_init(p_full, 256);
        mbar_init(dq_full, 1);
        fence_mbar_init();
        prefetch_tmap(&tmQ_hi); prefetch_tmap(&tmQ_lo); prefetch_tmap(&tmK_hi); prefetch_tmap(&tmK_lo);
        prefetch_tmap(&tmV_hi); prefetch_tmap(&tmV_lo); prefetch_tmap(&tmO_hi); prefetch_tmap(&tmO_lo);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            mbar_expect_tx(kv_full, 4 * AB_PLANE);
            tma_load_3d(smem + AB_SM_K, &tmK_hi, kv_full, p.k_off, k0, b);
            tma_load_3d(smem + AB_SM_K + AB_PLANE, &tmK_lo, kv_full, p.k_off, k0, b);
            tma_load_3d(smem + AB_SM_V, &tmV_hi, kv_full, p.v_off, k0, b);
            tma_load_3d(smem + AB_SM_V + AB_PLANE, &tmV_lo, kv_full, p.v_off, k0, b);
            for (int i = 0; i < p.nq; ++i) {
                if (i > 0) mbar_wait(dq_full, (uint32_t)((i - 1) & 1));       // the MMAs that read Q_{i-1}, dO_{i-1} have retired
                const int q0 = i * AB_BQ;
                mbar_expect_tx(qd_full, 4 * AB_PLANE + 2 * AB_BQ * 4);
                tma_load_3d(smem + AB_SM_Q, &tmQ_hi, qd_full, p.q_off, q0, b);
                tma_load_3d(smem + AB_SM_Q + AB_PLANE, &tmQ_lo, qd_full, p.q_off, q0, b);
                tma_load_3d(smem + AB_SM_DO, &tmO_hi, qd_full, p.do_off, q0, b);
                tma_load_3d(smem + AB_SM_DO + AB_PLANE, &tmO_lo, qd_full, p.do_off, q0, b);
                bulk_g2s(s_lse, p.lse + (size_t)b * p.s + q0, AB_BQ * 4, qd_full);
                bulk_g2s(s_rd, p.rowdot + (size_t)b * p.s + q0, AB_BQ * 4, qd_full);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (one thread) =================
        if (elect_one()) {
            // D = f32, A = B = bf16; N, M = 128.  bit 15: A MN-major, bit 16: B MN-major
            const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t idesc_ts = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(AB_D >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t idesc_dq = idesc_ts | (1u << 15);
            const uint32_t sK = smem_u32(smem + AB_SM_K), sV = smem_u32(smem + AB_SM_V), sQ = smem_u32(smem + AB_SM_Q);
            const uint32_t sO = smem_u32(smem + AB_SM_DO), sS = smem_u32(smem + AB_SM_DS);
            mbar_wait(kv_full, 0);
            for (int i = 0; i < p.nq; ++i) {
                mbar_wait(qd_full, (uint32_t)(i & 1));
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < AB_D / 16; ++k) {          // S^T = K_j Q_i^T,  dP^T = V_j dO_i^T   (contraction over the 64 channels)
                    const uint64_t koff = (uint64_t)(k * 2);
                    const uint64_t dk_hi = make_smem_desc(sK) + koff, dk_lo = make_smem_desc(sK + AB_PLANE) + koff;
                    const uint64_t dq_hi = make_smem_desc(sQ) + koff, dq_lo = make_smem_desc(sQ + AB_PLANE) + koff;
                    umma_bf16(tmem_base + AB_T_S, dk_lo, dq_hi, idesc_s, k != 0);
                    umma_bf16(tmem_base + AB_T_S, dk_hi, dq_lo, idesc_s, 1u);
                    umma_bf16(tmem_base + AB_T_S, dk_hi, dq_hi, idesc_s, 1u);
                }
#pragma unroll
                for (int k = 0; k < AB_D / 16; ++k) {
                    const uint64_t koff = (uint64_t)(k * 2);
                    const uint64_t dv_hi = make_smem_desc(sV) + koff, dv_lo = make_smem_desc(sV + AB_PLANE) + koff;
                    const uint64_t do_hi = make_smem_desc(sO) + koff, do_lo = make_smem_desc(sO + AB_PLANE) + koff;
                    umma_bf16(tmem_base + AB_T_DP, dv_lo, do_hi, idesc_s, k != 0);
                    umma_bf16(tmem_base + AB_T_DP, dv_hi, do_lo, idesc_s, 1u);
                    umma_bf16(tmem_base + AB_T_DP, dv_hi, do_hi, idesc_s, 1u);
                }
                tc_commit(s_full);
                mbar_wait(p_full, (uint32_t)(i & 1));
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < AB_BQ / 16; ++ks) {      // dV_j += P^T dO_i,  dK_j += dS^T Q_i   (contraction over the 128 queries)
                    // packed A columns: every 32-query chunk c keeps its hi pairs at [32c, 32c+16) and its lo pairs at [32c+16, 32c+32)
                    const uint32_t acol = (uint32_t)(32 * (ks >> 1) + (ks & 1) * 8);
                    const uint32_t row_off = (uint32_t)(ks * 16 * 128);           // 16 query rows of the MN-major B tiles
                    const uint64_t bo_hi = ab_desc_mn(sO + row_off, 0), bo_lo = ab_desc_mn(sO + AB_PLANE + row_off, 0);
                    const uint64_t bq_hi = ab_desc_mn(sQ + row_off, 0), bq_lo = ab_desc_mn(sQ + AB_PLANE + row_off, 0);
                    const uint32_t p_hi = tmem_base + AB_T_S + acol, p_lo = p_hi + 16;
                    const uint32_t s_hi = tmem_base + AB_T_DP + acol, s_lo = s_hi + 16;
                    umma_bf16_ts(tmem_base + AB_T_DV, p_lo, bo_hi, idesc_ts, (i | ks) != 0);
                    umma_bf16_ts(tmem_base + AB_T_DV, p_hi, bo_lo, idesc_ts, 1u);
                    umma_bf16_ts(tmem_base + AB_T_DV, p_hi, bo_hi, idesc_ts, 1u);
                    umma_bf16_ts(tmem_base + AB_T_DK, s_lo, bq_hi, idesc_ts, (i | ks) != 0);
                    umma_bf16_ts(tmem_base + AB_T_DK, s_hi, bq_lo, idesc_ts, 1u);
                    umma_bf16_ts(tmem_base + AB_T_DK, s_hi, bq_hi, idesc_ts, 1u);
                }
#pragma unroll
                for (int ks = 0; ks < AB_BK / 16; ++ks) {      // dQ_i = dS K_j   (contraction over the 128 keys; both operands MN-major)
                    const uint32_t row_off = (uint32_t)(ks * 16 * 128);
                    const uint64_t a_hi = ab_desc_mn(sS + row_off, AB_PLANE), a_lo = ab_desc_mn(sS + 2 * AB_PLANE + row_off, AB_PLANE);
                    const uint64_t bk_hi = ab_desc_mn(sK + row_off, 0), bk_lo = ab_desc_mn(sK + AB_PLANE + row_off, 0);
                    umma_bf16(tmem_base + AB_T_DQ, a_lo, bk_hi, idesc_dq, ks != 0);
                    umma_bf16(tmem_base + AB_T_DQ, a_hi, bk_lo, idesc_dq, 1u);
                    umma_bf16(tmem_base + AB_T_DQ, a_hi, bk_hi, idesc_dq, 1u);
                }
                tc_commit(dq_full);
            }
        }
    } else {
        // ================= element-wise warps: thread <-> (key row, half of the query columns) =================
        const int q4 = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = q4 * 32 + lane;                                     // key row within the block = TMEM lane
        const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
        const float kLog2e = 1.4426950408889634f;
        const uint32_t xr = (uint32_t)(row & 7);
        // dS^T tile: atom `half` (64 queries) of [128 key rows][128 B], hi plane then lo plane
        const uint32_t ds_hi = smem_u32(smem + AB_SM_DS) + (uint32_t)half * AB_PLANE + (uint32_t)row * 128;
        const uint32_t ds_lo = ds_hi + 2 * AB_PLANE;
        for (int i = 0; i < p.nq; ++i) {
            mbar_wait(qd_full, (uint32_t)(i & 1));                          // lse_i / rowdot_i are in shared memory
            mbar_wait(s_full, (uint32_t)(i & 1));
            tc_fence_after();
            const uint32_t ts = tmem_base + AB_T_S + (uint32_t)(half * 64) + lane_sel;
            const uint32_t td = tmem_base + AB_T_DP + (uint32_t)(half * 64) + lane_sel;
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {                                   // 32 queries at a time: outputs stay inside the chunk's own columns
                uint32_t sv[32], dv[32];
                tmem_ld32_nowait(ts + (uint32_t)(c * 32), sv);
                tmem_ld32_nowait(td + (uint32_t)(c * 32), dv);
                tmem_ld_wait();
                uint32_t ph[16], pl[16], sh[16], sl[16];                    // packed bf16 pairs: 32 queries -> 16 columns per plane
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int qa = half * 64 + c * 32 + 2 * j;
                    const float la = s_lse[qa] * kLog2e, lb = s_lse[qa + 1] * kLog2e;
                    float pa, pb;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pa) : "f"(fmaf(__uint_as_float(sv[2 * j]), kLog2e, -la)));
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pb) : "f"(fmaf(__uint_as_float(sv[2 * j + 1]), kLog2e, -lb)));
                    const float da = pa * (__uint_as_float(dv[2 * j]) - s_rd[qa]);
                    const float db = pb * (__uint_as_float(dv[2 * j + 1]) - s_rd[qa + 1]);
                    split2(pa, pb, ph[j], pl[j]);
                    split2(da, db, sh[j], sl[j]);
                }
                // dS^T -> shared memory (MN-major A operand of dQ = dS K): 32 queries = 4 chunks of 16 B per plane, 128B-swizzled by key row
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint32_t off = (((uint32_t)(c * 4 + g) ^ xr) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ds_hi + off), "r"(sh[4 * g]), "r"(sh[4 * g + 1]), "r"(sh[4 * g + 2]),
                                 "r"(sh[4 * g + 3]) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ds_lo + off), "r"(sl[4 * g]), "r"(sl[4 * g + 1]), "r"(sl[4 * g + 2]),
                                 "r"(sl[4 * g + 3]) : "memory");
                }
                // P^T and dS^T packed, in place over the chunk's own S^T / dP^T columns: hi at [32c', +16), lo at [32c' + 16, +16)
                tmem_st16(ts + (uint32_t)(c * 32), ph);
                tmem_st16(ts + (uint32_t)(c * 32 + 16), pl);
                tmem_st16(td + (uint32_t)(c * 32), sh);
                tmem_st16(td + (uint32_t)(c * 32 + 16), sl);
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(p_full);
            // dQ_i (lanes = query rows) -> global: this warp adds 32 of the 64 channels of its 32 rows
            mbar_wait(dq_full, (uint32_t)(i & 1));
            tc_fence_after();
            uint32_t acc[32];
            tmem_ld32(tmem_base + AB_T_DQ + (uint32_t)(half * 32) + lane_sel, acc);
            float* dst = p.dq + ((size_t)b * p.s + (size_t)i * AB_BQ + row) * p.dq_ld + p.dq_off + half * 32;
#pragma unroll
            for (int g = 0; g < 8; ++g)
                ab_red_add_v4(dst + 4 * g, __uint_as_float(acc[4 * g]), __uint_as_float(acc[4 * g + 1]), __uint_as_float(acc[4 * g + 2]),
                              __uint_as_float(acc[4 * g + 3]));
            tc_fence_before();
        }
        // dV_j, dK_j (lanes = key rows) -> global
        {
            uint32_t acc[32];
            const size_t krow = (size_t)b * p.s + k0 + row;
            tmem_ld32(tmem_base + AB_T_DV + (uint32_t)(half * 32) + lane_sel, acc);
            float* dv_dst = p.dv + krow * p.dv_ld + p.dv_off + half * 32;
#pragma unroll
            for (int g = 0; g < 8; ++g)
                ab_red_add_v4(dv_dst + 4 * g, __uint_as_float(acc[4 * g]), __uint_as_float(acc[4 * g + 1]), __uint_as_float(acc[4 * g + 2]),
                              __uint_as_float(acc[4 * g + 3]));
            tmem_ld32(tmem_base + AB_T_DK + (uint32_t)(half * 32) + lane_sel, acc);
            float* dk_dst = p.dk + krow * p.dk_ld + p.dk_off + half * 32;
#pragma unroll
            for (int g = 0; g < 8; ++g)
                ab_red_add_v4(dk_dst + 4 * g, __uint_as_float(acc[4 * g]), __uint_as_float(acc[4 * g + 1]), __uint_as_float(acc[4 * g + 2]),
                              __uint_as_float(acc[4 * g + 3]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------------
// Version 2 (default): the same formulation, software-pipelined over 64-QUERY sub-blocks within the same 193 KB of shared memory.
//   * Q_s / dO_s (+ lse, rowdot) are double-buffered: 2 x 32 KB in the region that held one 128-query block, so the TMA load of
//     sub-block s+2 runs while s and s+1 are processed (version 1 could start a load only after every MMA of the previous block had
//     retired, then waited out its full latency);
//   * S^T / dP^T are double-buffered in tensor memory (2 x 64 + 2 x 64 columns): the MMA thread issues S^T, dP^T of sub-block s+1
//     BEFORE it waits for the element-wise warps of sub-block s, so the tensor pipe works while P^T, dS^T of s are being formed;
//   * dV_j += P^T dO_s and dK_j += dS^T Q_s run per sub-block (K = 64 queries); dQ = dS K_j runs once per PAIR of sub-blocks with
//     M = 128 queries — the two 64-query atoms of the dS^T shared-memory tile are exactly the two sub-blocks of a pair.
// TMEM columns: S^T/P^T [0,64) [64,128), dP^T/dS^T [128,192) [192,256), dV [256,320), dK [320,384), dQ [384,448).
constexpr int A2_BQ = 64;                            // queries per sub-block
constexpr int A2_QPLANE = A2_BQ * 128;               // bytes of one [64 rows][64 bf16] plane
constexpr int A2_SM_QD = 4 * AB_PLANE;               // two buffers of {Q_hi, Q_lo, dO_hi, dO_lo}: 2 x 32 KB
constexpr int A2_T_S = 0, A2_T_DP = 128, A2_T_DV = 256, A2_T_DK = 320, A2_T_DQ = 384;

__global__ void __launch_bounds__(AB_THREADS, 1)
attention_bwd_kernel_v2(const __grid_constant__ CUtensorMap tmQ_hi, const __grid_constant__ CUtensorMap tmQ_lo,
                        const __grid_constant__ CUtensorMap tmK_hi, const __grid_constant__ CUtensorMap tmK_lo,
                        const __grid_constant__ CUtensorMap tmV_hi, const __grid_constant__ CUtensorMap tmV_lo,
                        const __grid_constant__ CUtensorMap tmO_hi, const __grid_constant__ CUtensorMap tmO_lo, const AttnBwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AB_SM_BAR);
    uint64_t* kv_full = bars;          // K_j, V_j resident
    uint64_t* qd_full = bars + 1;      // [2] Q_s, dO_s, lse_s, rowdot_s of buffer b loaded
    uint64_t* qd_free = bars + 3;      // [2] the MMAs that read buffer b (dV, dK of its sub-block) have retired
    uint64_t* s_full = bars + 5;       // [2] S^T and dP^T of buffer b in TMEM
    uint64_t* p_full = bars + 7;       // [2] P^T, dS^T of buffer b written (TMEM + shared memory): 256 arrivals
    uint64_t* dq_full = bars + 9;      // dQ of a pair complete (all earlier MMAs retired)
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 10);
    float* s_vec = reinterpret_cast<float*>(smem + AB_SM_VEC);     // [2 buffers][lse 64 | rowdot 64]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * AB_BK;
    const int b = blockIdx.y;
    const int nsub = p.s / A2_BQ;      // even: s is a multiple of 128

    if (threadIdx.x == 0) {
        mbar_init(kv_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&qd_full[i], 1);
            mbar_init(&qd_free[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 256);
        }
        mbar_init(dq_full, 1);
        fence_mbar_init();
        prefetch_tmap(&tmQ_hi); prefetch_tmap(&tmQ_lo); prefetch_tmap(&tmK_hi); prefetch_tmap(&tmK_lo);
        prefetch_tmap(&tmV_hi); prefetch_tmap(&tmV_lo); prefetch_tmap(&tmO_hi); prefetch_tmap(&tmO_lo);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            mbar_expect_tx(kv_full, 4 * AB_PLANE);
            tma_load_3d(smem + AB_SM_K, &tmK_hi, kv_full, p.k_off, k0, b);
            tma_load_3d(smem + AB_SM_K + AB_PLANE, &tmK_lo, kv_full, p.k_off, k0, b);
            tma_load_3d(smem + AB_SM_V, &tmV_hi, kv_full, p.v_off, k0, b);
            tma_load_3d(smem + AB_SM_V + AB_PLANE, &tmV_lo, kv_full, p.v_off, k0, b);
            for (int s = 0; s < nsub; ++s) {
                const int bf = s & 1;
                if (s >= 2) mbar_wait(&qd_free[bf], (uint32_t)(((s >> 1) - 1) & 1));
                const int q0 = s * A2_BQ;
                uint8_t* qd = smem + A2_SM_QD + bf * 4 * A2_QPLANE;
                mbar_expect_tx(&qd_full[bf], 4 * A2_QPLANE + 2 * A2_BQ * 4);
                tma_load_3d(qd, &tmQ_hi, &qd_full[bf], p.q_off, q0, b);
                tma_load_3d(qd + A2_QPLANE, &tmQ_lo, &qd_full[bf], p.q_off, q0, b);
                tma_load_3d(qd + 2 * A2_QPLANE, &tmO_hi, &qd_full[bf], p.do_off, q0, b);
                tma_load_3d(qd + 3 * A2_QPLANE, &tmO_lo, &qd_full[bf], p.do_off, q0, b);
                bulk_g2s(s_vec + bf * 2 * A2_BQ, p.lse + (size_t)b * p.s + q0, A2_BQ * 4, &qd_full[bf]);
                bulk_g2s(s_vec + bf * 2 * A2_BQ + A2_BQ, p.rowdot + (size_t)b * p.s + q0, A2_BQ * 4, &qd_full[bf]);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (one thread) =================
        if (elect_one()) {
            // D = f32, A = B = bf16, M = 128.  bit 15: A MN-major, bit 16: B MN-major
            const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(A2_BQ >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t idesc_ts = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(AB_D >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t idesc_dq = idesc_ts | (1u << 15);
            const uint32_t sK = smem_u32(smem + AB_SM_K), sV = smem_u32(smem + AB_SM_V), sS = smem_u32(smem + AB_SM_DS);
            auto issue_scores = [&](int s) {                 // S^T = K_j Q_s^T, dP^T = V_j dO_s^T into TMEM buffer s & 1
                const int bf = s & 1;
                mbar_wait(&qd_full[bf], (uint32_t)((s >> 1) & 1));
                tc_fence_after();
                const uint32_t sQ = smem_u32(smem + A2_SM_QD + bf * 4 * A2_QPLANE), sO = sQ + 2 * A2_QPLANE;
                const uint32_t tS = tmem_base + A2_T_S + (uint32_t)(bf * A2_BQ), tP = tmem_base + A2_T_DP + (uint32_t)(bf * A2_BQ);
#pragma unroll
                for (int k = 0; k < AB_D / 16; ++k) {
                    const uint64_t koff = (uint64_t)(k * 2);
                    const uint64_t dk_hi = make_smem_desc(sK) + koff, dk_lo = make_smem_desc(sK + AB_PLANE) + koff;
                    const uint64_t dq_hi = make_smem_desc(sQ) + koff, dq_lo = make_smem_desc(sQ + A2_QPLANE) + koff;
                    umma_bf16(tS, dk_lo, dq_hi, idesc_s, k != 0);
                    umma_bf16(tS, dk_hi, dq_lo, idesc_s, 1u);
                    umma_bf16(tS, dk_hi, dq_hi, idesc_s, 1u);
                }
#pragma unroll
                for (int k = 0; k < AB_D / 16; ++k) {
                    const uint64_t koff = (uint64_t)(k * 2);
                    const uint64_t dv_hi = make_smem_desc(sV) + koff, dv_lo = make_smem_desc(sV + AB_PLANE) + koff;
                    const uint64_t do_hi = make_smem_desc(sO) + koff, do_lo = make_smem_desc(sO + A2_QPLANE) + koff;
                    umma_bf16(tP, dv_lo, do_hi, idesc_s, k != 0);
                    umma_bf16(tP, dv_hi, do_lo, idesc_s, 1u);
                    umma_bf16(tP, dv_hi, do_hi, idesc_s, 1u);
                }
                tc_commit(&s_full[bf]);
            };
            mbar_wait(kv_full, 0);
            issue_scores(0);
            for (int s = 0; s < nsub; ++s) {
                const int bf = s & 1;
                if (s + 1 < nsub) issue_scores(s + 1);       // the tensor pipe works on s + 1 while the element-wise warps form P^T, dS^T of s
                mbar_wait(&p_full[bf], (uint32_t)((s >> 1) & 1));
                tc_fence_after();
                const uint32_t sQ = smem_u32(smem + A2_SM_QD + bf * 4 * A2_QPLANE), sO = sQ + 2 * A2_QPLANE;
                const uint32_t tS = tmem_base + A2_T_S + (uint32_t)(bf * A2_BQ), tP = tmem_base + A2_T_DP + (uint32_t)(bf * A2_BQ);
#pragma unroll
                for (int ks = 0; ks < A2_BQ / 16; ++ks) {    // dV_j += P^T dO_s,  dK_j += dS^T Q_s   (contraction over the 64 queries)
                    // packed A columns: every 32-query chunk c keeps its hi pairs at [32c, 32c+16) and its lo pairs at [32c+16, 32c+32)
                    const uint32_t acol = (uint32_t)(32 * (ks >> 1) + (ks & 1) * 8);
                    const uint32_t row_off = (uint32_t)(ks * 16 * 128);           // 16 query rows of the MN-major B tiles
                    const uint64_t bo_hi = ab_desc_mn(sO + row_off, 0), bo_lo = ab_desc_mn(sO + A2_QPLANE + row_off, 0);
                    const uint64_t bq_hi = ab_desc_mn(sQ + row_off, 0), bq_lo = ab_desc_mn(sQ + A2_QPLANE + row_off, 0);
                    const uint32_t p_hi = tS + acol, p_lo = p_hi + 16;
                    const uint32_t s_hi = tP + acol, s_lo = s_hi + 16;
                    umma_bf16_ts(tmem_base + A2_T_DV, p_lo, bo_hi, idesc_ts, (s | ks) != 0);
                    umma_bf16_ts(tmem_base + A2_T_DV, p_hi, bo_lo, idesc_ts, 1u);
                    umma_bf16_ts(tmem_base + A2_T_DV, p_hi, bo_hi, idesc_ts, 1u);
                    umma_bf16_ts(tmem_base + A2_T_DK, s_lo, bq_hi, idesc_ts, (s | ks) != 0);
                    umma_bf16_ts(tmem_base + A2_T_DK, s_hi, bq_lo, idesc_ts, 1u);
                    umma_bf16_ts(tmem_base + A2_T_DK, s_hi, bq_hi, idesc_ts, 1u);
                }
                tc_commit(&qd_free[bf]);                      // Q_s / dO_s may be overwritten once these MMAs have retired
                if (s & 1) {
#pragma unroll
                    for (int ks = 0; ks < AB_BK / 16; ++ks) {  // dQ_pair = dS K_j   (M = 128 queries = both atoms; contraction over the 128 keys)
                        const uint32_t row_off = (uint32_t)(ks * 16 * 128);
                        const uint64_t a_hi = ab_desc_mn(sS + row_off, AB_PLANE), a_lo = ab_desc_mn(sS + 2 * AB_PLANE + row_off, AB_PLANE);
                        const uint64_t bk_hi = ab_desc_mn(sK + row_off, 0), bk_lo = ab_desc_mn(sK + AB_PLANE + row_off, 0);
                        umma_bf16(tmem_base + A2_T_DQ, a_lo, bk_hi, idesc_dq, ks != 0);
                        umma_bf16(tmem_base + A2_T_DQ, a_hi, bk_lo, idesc_dq, 1u);
                        umma_bf16(tmem_base + A2_T_DQ, a_hi, bk_hi, idesc_dq, 1u);
                    }
                    tc_commit(dq_full);
                }
            }
        }
    } else {
        // ================= element-wise warps: thread <-> (key row, one 32-query chunk of the sub-block) =================
        const int q4 = warp & 3;
        const int half = (warp - 2) >> 2;                                   // which 32-query chunk of the 64-query sub-block
        const int row = q4 * 32 + lane;                                     // key row within the block = TMEM lane
        const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
        const float kLog2e = 1.4426950408889634f;
        const uint32_t xr = (uint32_t)(row & 7);
        // dQ of a pair (lanes = its 128 query rows) -> global: this warp adds 32 of the 64 channels of its 32 rows
        auto dq_epilogue = [&](int pair) {
            uint32_t acc[32];
            tmem_ld32(tmem_base + A2_T_DQ + (uint32_t)(half * 32) + lane_sel, acc);
            float* dst = p.dq + ((size_t)b * p.s + (size_t)pair * 128 + row) * p.dq_ld + p.dq_off + half * 32;
#pragma unroll
            for (int g = 0; g < 8; ++g)
                ab_red_add_v4(dst + 4 * g, __uint_as_float(acc[4 * g]), __uint_as_float(acc[4 * g + 1]), __uint_as_float(acc[4 * g + 2]),
                              __uint_as_float(acc[4 * g + 3]));
            tc_fence_before();
        };
        for (int s = 0; s < nsub; ++s) {
            const int bf = s & 1;
            const uint32_t par = (uint32_t)((s >> 1) & 1);
            mbar_wait(&qd_full[bf], par);                                   // lse_s / rowdot_s are in shared memory
            mbar_wait(&s_full[bf], par);
            tc_fence_after();
            const uint32_t v_lse = smem_u32(s_vec + bf * 2 * A2_BQ + half * 32), v_rd = v_lse + A2_BQ * 4;
            const uint32_t ts = tmem_base + A2_T_S + (uint32_t)(bf * A2_BQ + half * 32) + lane_sel;
            const uint32_t td = tmem_base + A2_T_DP + (uint32_t)(bf * A2_BQ + half * 32) + lane_sel;
            // dS^T tile of the pair: atom (s & 1) = this sub-block's 64 queries, [128 key rows][128 B], hi planes then lo planes
            const uint32_t ds_hi = smem_u32(smem + AB_SM_DS) + (uint32_t)bf * AB_PLANE + (uint32_t)row * 128;
            const uint32_t ds_lo = ds_hi + 2 * AB_PLANE;
            uint32_t sv[32], dv[32];
            tmem_ld32_nowait(ts, sv);
            tmem_ld32_nowait(td, dv);
            tmem_ld_wait();
            uint32_t ph[16], pl[16], sh[16], sl[16];                        // packed bf16 pairs: 32 queries -> 16 columns per plane
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {                                // four queries per step: one 16-byte read of lse and of rowdot
                const float4 l4 = lds_f4(v_lse + j4 * 16), r4 = lds_f4(v_rd + j4 * 16);
                const float ls[4] = {l4.x, l4.y, l4.z, l4.w}, rs[4] = {r4.x, r4.y, r4.z, r4.w};
                float pv[4], dsv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pv[e]) : "f"(fmaf(__uint_as_float(sv[4 * j4 + e]), kLog2e, -ls[e] * kLog2e)));
                    dsv[e] = pv[e] * (__uint_as_float(dv[4 * j4 + e]) - rs[e]);
                }
                split2(pv[0], pv[1], ph[2 * j4], pl[2 * j4]);
                split2(pv[2], pv[3], ph[2 * j4 + 1], pl[2 * j4 + 1]);
                split2(dsv[0], dsv[1], sh[2 * j4], sl[2 * j4]);
                split2(dsv[2], dsv[3], sh[2 * j4 + 1], sl[2 * j4 + 1]);
            }
            // The dS^T atoms of the previous pair are read by its dQ MMA: its completion (dq_full) is awaited only HERE, after this
            // sub-block's loads and arithmetic, so that MMA runs while the element-wise work of the next sub-block is under way.
            if (s >= 2 && !(s & 1)) {
                mbar_wait(dq_full, (uint32_t)(((s >> 1) - 1) & 1));
                tc_fence_after();
            }
            // dS^T -> shared memory (MN-major A operand of dQ = dS K): 32 queries = 4 chunks of 16 B per plane, 128B-swizzled by key row
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint32_t off = (((uint32_t)(half * 4 + g) ^ xr) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ds_hi + off), "r"(sh[4 * g]), "r"(sh[4 * g + 1]), "r"(sh[4 * g + 2]),
                             "r"(sh[4 * g + 3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ds_lo + off), "r"(sl[4 * g]), "r"(sl[4 * g + 1]), "r"(sl[4 * g + 2]),
                             "r"(sl[4 * g + 3]) : "memory");
            }
            // P^T and dS^T packed, in place over the chunk's own S^T / dP^T columns: hi at [0, 16), lo at [16, 32) of the chunk
            tmem_st32_halves_nowait(ts, ph, pl);
            tmem_st32_halves_nowait(td, sh, sl);
            tmem_st_wait();
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&p_full[bf]);
            if (s >= 2 && !(s & 1)) dq_epilogue((s >> 1) - 1);              // the previous pair's dQ (its MMA completed: awaited above)
        }
        mbar_wait(dq_full, (uint32_t)(((nsub >> 1) - 1) & 1));              // last pair
        tc_fence_after();
        dq_epilogue((nsub >> 1) - 1);
        // dV_j, dK_j (lanes = key rows) -> global: every MMA has retired once the last pair's dq_full completed (waited above)
        {
            uint32_t acc[32];
            const size_t krow = (size_t)b * p.s + k0 + row;
            tc_fence_after();
            tmem_ld32(tmem_base + A2_T_DV + (uint32_t)(half * 32) + lane_sel, acc);
            float* dv_dst = p.dv + krow * p.dv_ld + p.dv_off + half * 32;
#pragma unroll
            for (int g = 0; g < 8; ++g)
                ab_red_add_v4(dv_dst + 4 * g, __uint_as_float(acc[4 * g]), __uint_as_float(acc[4 * g + 1]), __uint_as_float(acc[4 * g + 2]),
                              __uint_as_float(acc[4 * g + 3]));
            tmem_ld32(tmem_base + A2_T_DK + (uint32_t)(half * 32) + lane_sel, acc);
            float* dk_dst = p.dk + krow * p.dk_ld + p.dk_off + half * 32;
#pragma unroll
            for (int g = 0; g < 8; ++g)
                ab_red_add_v4(dk_dst + 4 * g, __uint_as_float(acc[4 * g]), __uint_as_float(acc[4 * g + 1]), __uint_as_float(acc[4 * g + 2]),
                              __uint_as_float(acc[4 * g + 3]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

static int ab_rows_map(CUtensorMap* map, const void* base, int row_len, int rows, int batch, int box_rows = 128) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return HUPR_ERR_CUDA;
    cuuint64_t dims[3] = {(cuuint64_t)row_len, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)row_len * 2, (cuuint64_t)rows * row_len * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? HUPR_OK : HUPR_ERR_CUDA;
}

}  // namespace hupr

extern "C" int hupr_attention_bwd(const hupr_attn_bwd_desc* d, void* stream) {
    using namespace hupr;
    if (!d || !d->q_hi || !d->q_lo || !d->k_hi || !d->k_lo || !d->v_hi || !d->v_lo || !d->do_hi || !d->do_lo) return HUPR_ERR_BAD_ARG;
    if (!d->lse || !d->rowdot || !d->dq || !d->dk || !d->dv) return HUPR_ERR_BAD_ARG;
    if (d->batch <= 0 || d->batch > 65535 || d->s <= 0 || d->s % 128 || d->c != AB_D) return HUPR_ERR_BAD_ARG;
    const int lds[4] = {d->q_ld, d->k_ld, d->v_ld, d->do_ld}, offs[4] = {d->q_off, d->k_off, d->v_off, d->do_off};
    for (int i = 0; i < 4; ++i)
        if (lds[i] % 8 || offs[i] % 8 || offs[i] < 0 || offs[i] + AB_D > lds[i]) return HUPR_ERR_BAD_ARG;
    const int olds[3] = {d->dq_ld, d->dk_ld, d->dv_ld}, ooffs[3] = {d->dq_off, d->dk_off, d->dv_off};
    for (int i = 0; i < 3; ++i)
        if (olds[i] % 4 || ooffs[i] % 4 || ooffs[i] < 0 || ooffs[i] + AB_D > olds[i]) return HUPR_ERR_BAD_ARG;
    const uintptr_t align_or = (uintptr_t)d->q_hi | (uintptr_t)d->q_lo | (uintptr_t)d->k_hi | (uintptr_t)d->k_lo | (uintptr_t)d->v_hi | (uintptr_t)d->v_lo |
                               (uintptr_t)d->do_hi | (uintptr_t)d->do_lo | (uintptr_t)d->lse | (uintptr_t)d->rowdot | (uintptr_t)d->dq | (uintptr_t)d->dk |
                               (uintptr_t)d->dv;
    if (align_or & 15) return HUPR_ERR_ALIGNMENT;
    if (int arch_rc = device_check_sm100()) return arch_rc;
    static bool configured[kMaxDevices] = {}, configured2[kMaxDevices] = {};
    if (int crc = ensure_smem_optin(attention_bwd_kernel, AB_SMEM, configured)) return crc;
    if (int crc = ensure_smem_optin(attention_bwd_kernel_v2, AB_SMEM, configured2)) return crc;
    static const bool v1 = getenv("HUPR_ATTN_BWD_V1") != nullptr;      // A/B switch: the unpipelined 128-query version
    const int qrows = v1 ? AB_BQ : A2_BQ;                              // rows of a Q / dO box
    CUtensorMap q_hi, q_lo, k_hi, k_lo, v_hi, v_lo, o_hi, o_lo;
    int rc;
    if ((rc = ab_rows_map(&q_hi, d->q_hi, d->q_ld, d->s, d->batch, qrows)) != HUPR_OK) return rc;
    if ((rc = ab_rows_map(&q_lo, d->q_lo, d->q_ld, d->s, d->batch, qrows)) != HUPR_OK) return rc;
    if ((rc = ab_rows_map(&k_hi, d->k_hi, d->k_ld, d->s, d->batch)) != HUPR_OK) return rc;
    if ((rc = ab_rows_map(&k_lo, d->k_lo, d->k_ld, d->s, d->batch)) != HUPR_OK) return rc;
    if ((rc = ab_rows_map(&v_hi, d->v_hi, d->v_ld, d->s, d->batch)) != HUPR_OK) return rc;
    if ((rc = ab_rows_map(&v_lo, d->v_lo, d->v_ld, d->s, d->batch)) != HUPR_OK) return rc;
    if ((rc = ab_rows_map(&o_hi, d->do_hi, d->do_ld, d->s, d->batch, qrows)) != HUPR_OK) return rc;
    if ((rc = ab_rows_map(&o_lo, d->do_lo, d->do_ld, d->s, d->batch, qrows)) != HUPR_OK) return rc;
    AttnBwdParams p;
    p.nq = d->s / AB_BQ; p.s = d->s;
    p.q_off = d->q_off; p.k_off = d->k_off; p.v_off = d->v_off; p.do_off = d->do_off;
    p.lse = d->lse; p.rowdot = d->rowdot;
    p.dq = d->dq; p.dq_ld = d->dq_ld; p.dq_off = d->dq_off;
    p.dk = d->dk; p.dk_ld = d->dk_ld; p.dk_off = d->dk_off;
    p.dv = d->dv; p.dv_ld = d->dv_ld; p.dv_off = d->dv_off;
    const dim3 grid(d->s / AB_BK, d->batch);
    if (v1) attention_bwd_kernel<<<grid, AB_THREADS, AB_SMEM, (cudaStream_t)stream>>>(q_hi, q_lo, k_hi, k_lo, v_hi, v_lo, o_hi, o_lo, p);
    else attention_bwd_kernel_v2<<<grid, AB_THREADS, AB_SMEM, (cudaStream_t)stream>>>(q_hi, q_lo, k_hi, k_lo, v_hi, v_lo, o_hi, o_lo, p);
    note_launches(1);
    return cudaGetLastError() == cudaSuccess ? HUPR_OK : HUPR_ERR_CUDA;
}
