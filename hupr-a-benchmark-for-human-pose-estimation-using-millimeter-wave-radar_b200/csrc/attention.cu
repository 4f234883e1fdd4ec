// Fused single-head spatial attention (flash-style) on tcgen05 + TMEM + TMA, sm_100a.
//
// Replaces MultiScaleCrossSelfAttentionPRGCN.attention (/root/reference/models/layers.py:126-133):
//   logits[n, m] = <Q[n, :], K[m, :]>   (NO 1/sqrt(d) scale)      n = query position, m = key position
//   P = softmax over the KEY axis m;   out[n, :] = sum_m P[n, m] V[m, :]  (+ V[n, :] on the cross branches, :146,148)
// The reference materialises the [S, S] logits (64 MiB fp32 per sample per attention at S = 4096); here they never leave
// the SM: one CTA owns 128 query rows and streams the keys in blocks of 128.
//
//   * operands are bf16 hi/lo split tensors; every contraction is three tensor-core products (hi*hi + lo*hi + hi*lo) with
//     fp32 accumulation in TMEM — fp32-equivalent, like the convolutions;
//   * warp 0: TMA producer (Q once, then a 2-stage ring of K / V^T blocks); warp 1: single-thread tcgen05.mma issuer;
//     warps 2..9: softmax — two warps per TMEM lane quarter; a thread owns one query row (TMEM lane) and half of the
//     block's key columns, so max / sum are thread-local up to one smem exchange with its partner warp;
//   * per key block: S = Q K^T -> TMEM (two S buffers: Q K^T runs two key blocks ahead, so the tensor pipe works on
//     P_{j-1} V_{j-1} and Q K_{j+1}^T while the softmax warps exponentiate block j); softmax warps read S (tcgen05.ld), update
//     the running max (exact: when a row's max grows, its O row is rescaled in TMEM once P_{j-1} V_{j-1} has retired), write
//     P = exp(S - max) as hi/lo bf16 into a 128B-swizzled smem tile; O += P V accumulates in TMEM; the epilogue divides by
//     the running sum, adds the residual and stores hi/lo bf16.
#include "tc.cuh"
#include "split.cuh"

#ifndef HUPR_ATTN_P_IN_TMEM
#define HUPR_ATTN_P_IN_TMEM 1      // 1: softmax probabilities stay in TMEM (tcgen05.st, A-from-TMEM MMA); 0: 128B-swizzled smem tile
#endif

namespace hupr {

constexpr int AT_BM = 128;       // queries per CTA
constexpr int AT_THREADS = 320;  // TMA warp, MMA warp, 8 softmax warps (two per TMEM lane quarter: each owns half of the columns)
constexpr int AT_SMEM = 232448;  // the 227 KiB per-CTA maximum: layout end + 896 B of alignment slack

// D = head dim = channels, BKV = keys per block.  (64, 128) serves level 1 (S = 4096), (128, 64) level 2 (S = 1024) of
// mscsa_prgcn.yaml; both fill the same 224 KiB of shared memory.
template <int D, int BKV>
struct AttnCfg {
    static constexpr int kQPlane = (D / 64) * AT_BM * 128;        // Q: D/64 swizzle atoms of [128 rows][128 B]
    static constexpr int kKPlane = (D / 64) * BKV * 128;          // K block: D/64 atoms of [BKV rows][128 B]
    static constexpr int kVPlane = (BKV / 64) * D * 128;          // V^T block: BKV/64 atoms of [D rows][128 B]
    static constexpr int kPPlane = (BKV / 64) * AT_BM * 128;      // P: BKV/64 atoms of [128 rows][128 B]
    static constexpr int kSmQ = 0;                                // Q_hi, Q_lo
    static constexpr int kSmKV = 2 * kQPlane;                     // 2 stages x {K_hi, K_lo, Vt_hi, Vt_lo}
    static constexpr int kStage = 2 * kKPlane + 2 * kVPlane;
    static constexpr int kSmP = kSmKV + 2 * kStage;               // P_hi, P_lo
    static constexpr int kSmBar = kSmP + 2 * kPPlane;
    static constexpr int kSmX = kSmBar + 128;                     // float [2 parity][2 halves][128 rows] row-max exchange
    static constexpr int kSmEnd = kSmX + 2048;
    static constexpr int kTmemO = 2 * BKV;                        // S0 = [0, BKV), S1 = [BKV, 2 BKV), O = [2 BKV, 2 BKV + D)
    static constexpr int kTmemP = 2 * BKV + D;                    // P_hi = [kTmemP, + BKV/2), P_lo = [.. + BKV/2, + BKV): bf16 pairs per column
    static constexpr int kTmemCols = (2 * BKV + D + BKV) <= 256 ? 256 : 512;
    static_assert(kSmEnd + 896 <= AT_SMEM, "attention shared-memory plan does not fit");
};

struct AttnParams {
    int nkv;                      // key blocks
    int q_off, k_off;             // first channel of the Q / K slices inside their rows
    const __nv_bfloat16* r_hi; const __nv_bfloat16* r_lo; int r_ld, r_off;
    __nv_bfloat16* o_hi; __nv_bfloat16* o_lo; int o_ld, o_off;
    int s;
    float* lse;                   // optional [batch][s]: log sum_m exp(logit[n, m]) per query row (saved for the backward pass)
};

__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

template <int D, int BKV>
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ_hi, const __grid_constant__ CUtensorMap tmQ_lo,
                 const __grid_constant__ CUtensorMap tmK_hi, const __grid_constant__ CUtensorMap tmK_lo,
                 const __grid_constant__ CUtensorMap tmV_hi, const __grid_constant__ CUtensorMap tmV_lo, const AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    using Cfg = AttnCfg<D, BKV>;
    if (smem + Cfg::kSmEnd > smem_raw + AT_SMEM) __trap();      // dynamic smem base less aligned than the 896-B slack allows
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kSmBar);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;      // [2]
    uint64_t* k_empty = bars + 3;     // [2]
    uint64_t* v_full = bars + 5;      // [2]
    uint64_t* v_empty = bars + 7;     // [2]
    uint64_t* s_full = bars + 9;      // [2]  S buffer (j & 1) holds Q K_j^T
    uint64_t* p_full = bars + 11;     // P_j written to smem (256 arrivals)
    uint64_t* pv_done = bars + 12;    // O += P_j V_j retired: O may be rescaled, the P tile may be overwritten
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * AT_BM;
    const int b = blockIdx.y;

    if (threadIdx.x == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
            mbar_init(&s_full[i], 1);
        }
        mbar_init(p_full, 256);
        mbar_init(pv_done, 1);
        fence_mbar_init();
        prefetch_tmap(&tmQ_hi); prefetch_tmap(&tmQ_lo); prefetch_tmap(&tmK_hi);
        prefetch_tmap(&tmK_lo); prefetch_tmap(&tmV_hi); prefetch_tmap(&tmV_lo);
    }
    if (warp == 1) {   // TMEM: S0 = columns [0,BKV), S1 = [BKV,2BKV), O = [2BKV, 2BKV+D)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(Cfg::kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t tmem_O = tmem_base + Cfg::kTmemO;
    pdl_wait();                  // programmatic dependent launch: the predecessor's writes are visible from here on (common.cuh)
    pdl_launch_dependents();     // the successor may start its prologue now; it waits the same way before touching memory

    if (warp == 0) {
        // ================= TMA producer: K ring and V ring are independent (K_j is free as soon as Q K_j^T retires) =================
        if (elect_one()) {
            mbar_expect_tx(q_full, 2 * Cfg::kQPlane);
#pragma unroll
            for (int a = 0; a < D / 64; ++a) {
                tma_load_3d(smem + Cfg::kSmQ + a * (AT_BM * 128), &tmQ_hi, q_full, p.q_off + 64 * a, q0, b);
                tma_load_3d(smem + Cfg::kSmQ + Cfg::kQPlane + a * (AT_BM * 128), &tmQ_lo, q_full, p.q_off + 64 * a, q0, b);
            }
            for (int j = 0; j < p.nkv; ++j) {
                const int s = j & 1;
                const uint32_t ph = (uint32_t)(((j >> 1) & 1) ^ 1);
                uint8_t* st = smem + Cfg::kSmKV + s * Cfg::kStage;
                const int k0 = j * BKV;
                mbar_wait(&k_empty[s], ph);
                mbar_expect_tx(&k_full[s], 2 * Cfg::kKPlane);
#pragma unroll
                for (int a = 0; a < D / 64; ++a) {
                    tma_load_3d(st + a * (BKV * 128), &tmK_hi, &k_full[s], p.k_off + 64 * a, k0, b);
                    tma_load_3d(st + Cfg::kKPlane + a * (BKV * 128), &tmK_lo, &k_full[s], p.k_off + 64 * a, k0, b);
                }
                mbar_wait(&v_empty[s], ph);
                mbar_expect_tx(&v_full[s], 2 * Cfg::kVPlane);
#pragma unroll
                for (int a = 0; a < BKV / 64; ++a) {      // 64-key atoms of the V^T block
                    tma_load_3d(st + 2 * Cfg::kKPlane + a * (D * 128), &tmV_hi, &v_full[s], k0 + 64 * a, 0, b);
                    tma_load_3d(st + 2 * Cfg::kKPlane + Cfg::kVPlane + a * (D * 128), &tmV_lo, &v_full[s], k0 + 64 * a, 0, b);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (one thread): Q K^T runs two key blocks ahead of P V =================
        if (elect_one()) {
            const uint32_t idesc_qk = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BKV >> 3) << 17) | ((uint32_t)(AT_BM >> 4) << 24);
            const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(AT_BM >> 4) << 24);
            const uint32_t sQ = smem_u32(smem + Cfg::kSmQ), sP = smem_u32(smem + Cfg::kSmP);
            auto issue_qk = [&](int j) {      // S[j & 1] = Q . K_j^T
                const int s = j & 1;
                mbar_wait(&k_full[s], (uint32_t)((j >> 1) & 1));
                tc_fence_after();
                const uint32_t st = smem_u32(smem + Cfg::kSmKV + s * Cfg::kStage);
                const uint32_t tS = tmem_base + (uint32_t)(s * BKV);
#pragma unroll
                for (int k = 0; k < D / 16; ++k) {
                    const uint32_t atom = (uint32_t)(k >> 2);
                    const uint64_t koff = (uint64_t)((k & 3) * 2);
                    const uint64_t dq_hi = make_smem_desc(sQ + atom * (AT_BM * 128)) + koff;
                    const uint64_t dq_lo = make_smem_desc(sQ + Cfg::kQPlane + atom * (AT_BM * 128)) + koff;
                    const uint64_t dk_hi = make_smem_desc(st + atom * (BKV * 128)) + koff;
                    const uint64_t dk_lo = make_smem_desc(st + Cfg::kKPlane + atom * (BKV * 128)) + koff;
                    umma_bf16(tS, dq_lo, dk_hi, idesc_qk, k != 0);
                    umma_bf16(tS, dq_hi, dk_lo, idesc_qk, 1u);
                    umma_bf16(tS, dq_hi, dk_hi, idesc_qk, 1u);
                }
                tc_commit(&s_full[s]);
                tc_commit(&k_empty[s]);
            };
            mbar_wait(q_full, 0);
            issue_qk(0);
            if (p.nkv > 1) issue_qk(1);
            for (int j = 0; j < p.nkv; ++j) {
                const int s = j & 1;
                mbar_wait(&v_full[s], (uint32_t)((j >> 1) & 1));
                mbar_wait(p_full, (uint32_t)(j & 1));
                tc_fence_after();
                const uint32_t st = smem_u32(smem + Cfg::kSmKV + s * Cfg::kStage);
#pragma unroll
                for (int k = 0; k < BKV / 16; ++k) {      // O (+)= P_j . V_j
                    const uint32_t atom = (uint32_t)(k >> 2);
                    const uint64_t koff = (uint64_t)((k & 3) * 2);
                    const uint64_t dv_hi = make_smem_desc(st + 2 * Cfg::kKPlane + atom * (D * 128)) + koff;
                    const uint64_t dv_lo = make_smem_desc(st + 2 * Cfg::kKPlane + Cfg::kVPlane + atom * (D * 128)) + koff;
#if HUPR_ATTN_P_IN_TMEM
                    // P is the A operand straight from tensor memory: 16 keys = 8 packed columns per k-step
                    const uint32_t ap_hi = tmem_base + (uint32_t)(Cfg::kTmemP + k * 8);
                    const uint32_t ap_lo = ap_hi + (uint32_t)(BKV / 2);
                    umma_bf16_ts(tmem_O, ap_lo, dv_hi, idesc_pv, (j | k) != 0);
                    umma_bf16_ts(tmem_O, ap_hi, dv_lo, idesc_pv, 1u);
                    umma_bf16_ts(tmem_O, ap_hi, dv_hi, idesc_pv, 1u);
#else
                    const uint64_t dp_hi = make_smem_desc(sP + atom * (AT_BM * 128)) + koff;
                    const uint64_t dp_lo = make_smem_desc(sP + Cfg::kPPlane + atom * (AT_BM * 128)) + koff;
                    umma_bf16(tmem_O, dp_lo, dv_hi, idesc_pv, (j | k) != 0);
                    umma_bf16(tmem_O, dp_hi, dv_lo, idesc_pv, 1u);
                    umma_bf16(tmem_O, dp_hi, dv_hi, idesc_pv, 1u);
#endif
                }
                tc_commit(&v_empty[s]);
                tc_commit(pv_done);
                if (j + 2 < p.nkv) issue_qk(j + 2);     // S[j & 1] was drained by the softmax warps before they signalled p_full(j)
            }
        }
    } else {
        // ================= softmax / epilogue warps: thread <-> (query row, column half) =================
        constexpr int SC = BKV / 2;              // S columns per thread
        constexpr int NV = SC / 32;              // 32-column TMEM loads per thread and block
        constexpr int OC = D / 2;                // O columns per thread
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;        // S columns [SC*half, +SC), O columns [OC*half, +OC)
        const int row = q * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        const uint32_t row_base = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
        const uint32_t sP_hi = smem_u32(smem + Cfg::kSmP) + row_base;
        const uint32_t sP_lo = smem_u32(smem + Cfg::kSmP + Cfg::kPPlane) + row_base;
        const uint32_t xch = smem_u32(smem + Cfg::kSmX);
        const uint32_t xr = (uint32_t)(row & 7);
        const float kLog2e = 1.4426950408889634f;
        float m_run = -INFINITY, nm_run = 0.f, l_run = 0.f;    // l_run: partial sum over this thread's column half; nm_run = -m_run*log2e
        for (int j = 0; j < p.nkv; ++j) {
            mbar_wait(&s_full[j & 1], (uint32_t)((j >> 1) & 1));
            tc_fence_after();
            uint32_t v[NV][32];
            const uint32_t tS = tmem_base + (uint32_t)((j & 1) * BKV + half * SC) + lane_sel;
#pragma unroll
            for (int c = 0; c < NV; ++c) tmem_ld32_nowait(tS + (uint32_t)(c * 32), v[c]);
            tmem_ld_wait();
            float mx = m_run;
#pragma unroll
            for (int c = 0; c < NV; ++c) {
#pragma unroll
                for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[c][i]));
            }
            // combine the row maximum with the warp that owns the other column half
            const uint32_t slot = xch + (uint32_t)(((j & 1) * 2) * 128 * 4);
            sts_f32(slot + (uint32_t)((half * 128 + row) * 4), mx);
            asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
            mx = fmaxf(mx, lds_f32(slot + (uint32_t)(((half ^ 1) * 128 + row) * 4)));
            const float nm = -mx * kLog2e;
            // P = exp(S - max) = 2^(S*log2e - max*log2e), split into hi/lo bf16 (kept in registers until the P tile is free)
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < NV; ++c) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float p0, p1;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(__uint_as_float(v[c][2 * i]), kLog2e, nm)));
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(__uint_as_float(v[c][2 * i + 1]), kLog2e, nm)));
                    sum += p0 + p1;
                    split2(p0, p1, v[c][2 * i], v[c][2 * i + 1]);       // v[2i] = hi pair, v[2i+1] = lo pair of columns (2i, 2i+1)
                }
            }
            if (j > 0) {
                mbar_wait(pv_done, (uint32_t)((j - 1) & 1));     // O += P_{j-1} V_{j-1} has retired
                tc_fence_after();
                const bool grew = mx > m_run;
                if (__any_sync(0xffffffffu, grew)) {     // exact online softmax: rescale this row of O when its max grows
                    const float alpha = exp2f(nm - nm_run);          // same rounded exponent offsets as the P values already summed
#pragma unroll 1
                    for (int c = 0; c < OC / 32; ++c) {
                        uint32_t o[32];
                        tmem_ld32(tmem_O + lane_sel + (uint32_t)(half * OC + c * 32), o);
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st32(tmem_O + lane_sel + (uint32_t)(half * OC + c * 32), o);
                    }
                    l_run *= alpha;
                }
            }
            m_run = mx;
            nm_run = nm;
            l_run += sum;
#if HUPR_ATTN_P_IN_TMEM
            {   // P stays in tensor memory (A operand of the P.V MMAs): this thread's SC keys are SC/2 packed columns of each plane
                uint32_t ph[NV * 16], pl[NV * 16];
#pragma unroll
                for (int c = 0; c < NV; ++c) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) { ph[c * 16 + i] = v[c][2 * i]; pl[c * 16 + i] = v[c][2 * i + 1]; }
                }
                const uint32_t tp = tmem_base + (uint32_t)(Cfg::kTmemP + half * (SC / 2)) + lane_sel;
                if constexpr (NV == 2) {
                    tmem_st32(tp, ph);
                    tmem_st32(tp + (uint32_t)(BKV / 2), pl);
                } else {
                    tmem_st16(tp, ph);
                    tmem_st16(tp + (uint32_t)(BKV / 2), pl);
                }
            }
#else
            // 128B-swizzled K-major P tile: key column col of this row lives in atom col/64, 16-byte chunk ((col%64)/8) ^ (row & 7)
#pragma unroll
            for (int c = 0; c < NV; ++c) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int col = half * SC + c * 32 + g * 8;
                    const uint32_t off = (uint32_t)((col >> 6) * (AT_BM * 128)) + ((((uint32_t)(col & 63) >> 3) ^ xr) << 4);
                    sts_v4(sP_hi + off, v[c][8 * g], v[c][8 * g + 2], v[c][8 * g + 4], v[c][8 * g + 6]);
                    sts_v4(sP_lo + off, v[c][8 * g + 1], v[c][8 * g + 3], v[c][8 * g + 5], v[c][8 * g + 7]);
                }
            }
            fence_proxy_async();     // generic-proxy smem writes -> visible to the tensor core (async proxy)
#endif
            tc_fence_before();
            mbar_arrive(p_full);
        }
        // epilogue: O / l (+ residual) -> hi/lo bf16; the two halves first add up their partial sums
        const uint32_t lslot = xch + (uint32_t)(((p.nkv & 1) * 2) * 128 * 4);
        sts_f32(lslot + (uint32_t)((half * 128 + row) * 4), l_run);
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
        const float l_tot = l_run + lds_f32(lslot + (uint32_t)(((half ^ 1) * 128 + row) * 4));
        const float inv = 1.0f / l_tot;
        mbar_wait(pv_done, (uint32_t)((p.nkv - 1) & 1));
        tc_fence_after();
        const size_t pos = (size_t)b * p.s + q0 + row;
        if (p.lse && half == 0) p.lse[pos] = m_run + __logf(l_tot);
#pragma unroll 1
        for (int c = 0; c < OC / 32; ++c) {
            uint32_t v[32];
            tmem_ld32(tmem_O + lane_sel + (uint32_t)(half * OC + c * 32), v);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = __uint_as_float(v[g * 8 + i]) * inv;
                const int ch = half * OC + c * 32 + g * 8;
                if (p.r_hi) {
                    float r[8];
                    const size_t roff = pos * p.r_ld + p.r_off + ch;
                    load8(p.r_hi + roff, p.r_lo ? p.r_lo + roff : nullptr, r);
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] += r[i];
                }
                const size_t ooff = pos * p.o_ld + p.o_off + ch;
                store8(p.o_hi + ooff, p.o_lo + ooff, o);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::kTmemCols));
    }
}

static int encode_rows_map(CUtensorMap* map, const void* base, int row_len, int rows, int batch, int box_rows) {
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

template <int D, int BKV>
static int launch_attention(const hupr_attn_desc* d, cudaStream_t stream) {
    static bool configured[kMaxDevices] = {};
    if (int crc = ensure_smem_optin(attention_kernel<D, BKV>, AT_SMEM, configured)) return crc;
    CUtensorMap q_hi, q_lo, k_hi, k_lo, v_hi, v_lo;
    int rc;
    if ((rc = encode_rows_map(&q_hi, d->q_hi, d->q_ld, d->s, d->batch, AT_BM)) != HUPR_OK) return rc;
    if ((rc = encode_rows_map(&q_lo, d->q_lo, d->q_ld, d->s, d->batch, AT_BM)) != HUPR_OK) return rc;
    if ((rc = encode_rows_map(&k_hi, d->k_hi, d->k_ld, d->s, d->batch, BKV)) != HUPR_OK) return rc;
    if ((rc = encode_rows_map(&k_lo, d->k_lo, d->k_ld, d->s, d->batch, BKV)) != HUPR_OK) return rc;
    if ((rc = encode_rows_map(&v_hi, d->vt_hi, d->s, D, d->batch, D)) != HUPR_OK) return rc;
    if ((rc = encode_rows_map(&v_lo, d->vt_lo, d->s, D, d->batch, D)) != HUPR_OK) return rc;
    AttnParams p;
    p.nkv = d->s / BKV;
    p.q_off = d->q_off; p.k_off = d->k_off;
    p.r_hi = (const __nv_bfloat16*)d->r_hi; p.r_lo = (const __nv_bfloat16*)d->r_lo; p.r_ld = d->r_ld; p.r_off = d->r_off;
    p.o_hi = (__nv_bfloat16*)d->o_hi; p.o_lo = (__nv_bfloat16*)d->o_lo; p.o_ld = d->o_ld; p.o_off = d->o_off;
    p.s = d->s;
    p.lse = d->lse;
    const dim3 grid(d->s / AT_BM, d->batch);
    launch_k(attention_kernel<D, BKV>, grid, dim3(AT_THREADS), (size_t)AT_SMEM, stream, q_hi, q_lo, k_hi, k_lo, v_hi, v_lo, p);
    note_launches(1);
    return cudaGetLastError() == cudaSuccess ? HUPR_OK : HUPR_ERR_CUDA;
}

}  // namespace hupr

extern "C" int hupr_attention_fwd(const hupr_attn_desc* d, void* stream) {
    using namespace hupr;
    if (!d || !d->q_hi || !d->q_lo || !d->k_hi || !d->k_lo || !d->vt_hi || !d->vt_lo || !d->o_hi || !d->o_lo) return HUPR_ERR_BAD_ARG;
    if (d->batch <= 0 || d->batch > 65535 || d->s <= 0 || d->s % AT_BM || (d->c != 64 && d->c != 128)) return HUPR_ERR_BAD_ARG;
    const int c = d->c;
    if (d->q_ld % 8 || d->k_ld % 8 || d->q_off % 8 || d->k_off % 8 || d->q_off + c > d->q_ld || d->k_off + c > d->k_ld) return HUPR_ERR_BAD_ARG;
    if (d->o_ld % 8 || d->o_off % 8 || d->o_off + c > d->o_ld) return HUPR_ERR_BAD_ARG;
    if (d->r_hi && (d->r_ld % 8 || d->r_off % 8 || d->r_off + c > d->r_ld)) return HUPR_ERR_BAD_ARG;
    const uintptr_t align_or = (uintptr_t)d->q_hi | (uintptr_t)d->q_lo | (uintptr_t)d->k_hi | (uintptr_t)d->k_lo | (uintptr_t)d->vt_hi |
                               (uintptr_t)d->vt_lo | (uintptr_t)d->o_hi | (uintptr_t)d->o_lo | (uintptr_t)d->r_hi | (uintptr_t)d->r_lo;
    if (align_or & 15) return HUPR_ERR_ALIGNMENT;
    if (int arch_rc = device_check_sm100()) return arch_rc;
    return c == 64 ? launch_attention<64, 128>(d, (cudaStream_t)stream) : launch_attention<128, 64>(d, (cudaStream_t)stream);
}
