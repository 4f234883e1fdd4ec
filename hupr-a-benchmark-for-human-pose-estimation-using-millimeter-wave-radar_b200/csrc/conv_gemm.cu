// Implicit-GEMM convolution / GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// One kernel serves every dense contraction of MSCSA-PRGCN (reference models/layers.py):
//   Conv3d 3x3x3 (layers.py:45-63,195-206), temporal-merge Conv3d (T,1,1) (:208-210), Conv2d 3x3 (:24-32,81-95),
//   1x1 q/k projections (:116-123), the head conv (:94) and both attention matmuls (:126-133).
//
// Formulation:  D[pos, co] = sum_{tap} sum_{ci} A[pos + off(tap), ci] * W[tap][co][ci]
//   * activations are channels-last ([N][D][H][W][C]) bf16; fp32-equivalent accuracy comes from a hi/lo split
//     (x = hi + lo, both bf16) and three tensor-core products per k-step (hi*hi + lo*hi + hi*lo), fp32 accumulate in TMEM;
//   * no im2col buffer: for every filter tap the TMA engine loads the shifted [BH x BW] x 64-channel box straight from
//     the activation tensor (5-D tiled tensor map, SWIZZLE_128B, out-of-bounds = zero fill gives the conv padding);
//   * warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread tcgen05.mma issuer,
//     warps 2..9 = epilogue, two per TMEM lane quarter (each half of the tile's columns) (tcgen05.ld -> scale/shift (+residual) -> PReLU-style slope -> bf16 hi/lo or fp32 stores);
//   * mbarrier ring (full/empty) between TMA and MMA, tcgen05.commit releases stages and publishes the accumulator.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "hupr_internal.h"
#include "tc.cuh"
#include "conv_common.cuh"

namespace hupr {

constexpr int BM = 128;   // output positions per CTA tile (TMEM lanes)
constexpr int BK = 64;    // channels per k-block = one 128-byte swizzle row of bf16
constexpr int kConvThreads = 320;   // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter, each owning half of the tile's columns)

template <int BN, int NPROD, bool TSTORE = false>
struct ConvCfg {
    static constexpr int kPlanes = (NPROD == 3) ? 2 : 1;
    static constexpr int kABytes = BM * BK * 2;
    static constexpr int kBBytes = BN * BK * 2;
    static constexpr int kStageBytes = kPlanes * (kABytes + kBBytes);
    // TSTORE: the output tile is staged in shared memory ([hi | lo] x BN/64 blocks of [128 rows][128 B], 128B-swizzled) and written
    // with TMA stores instead of one-row-per-thread global stores
    static constexpr int kStoreBytes = TSTORE ? 2 * BM * BN * 2 : 0;
    static constexpr int kStagesMax = (200 * 1024 - kStoreBytes) / kStageBytes;
    static constexpr int kStages = kStagesMax > 6 ? 6 : kStagesMax;
    static constexpr int kSmemBytes = kStages * kStageBytes + kStoreBytes + 1024 /*align*/ + 256 /*barriers*/;
    static_assert(kStages >= 2, "conv smem plan");
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }


template <int BN, int NPROD, bool TSTORE>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                 const __grid_constant__ CUtensorMap tmO_hi, const __grid_constant__ CUtensorMap tmO_lo, const ConvParams p) {
    // Persistent: CTA b walks work items b, b + gridDim.x, ... (item = (split-K slice, column tile, 128-position tile)); the smem ring
    // runs continuously across items and two TMEM accumulators let the epilogue of one item overlap the main loop of the next.
    using Cfg = ConvCfg<BN, NPROD, TSTORE>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* stg = smem + Cfg::kStages * Cfg::kStageBytes;          // TSTORE: output staging (1024-B aligned: stage sizes are multiples of 1 KiB)
    uint64_t* bars = reinterpret_cast<uint64_t*>(stg + Cfg::kStoreBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + Cfg::kStages;
    uint64_t* accum_full = bars + 2 * Cfg::kStages;          // [2]
    uint64_t* accum_empty = bars + 2 * Cfg::kStages + 2;     // [2]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_kb = p.kd * p.kh * p.kw * p.cin_blocks;
    const int total_items = p.m_tiles * p.n_tiles * p.k_split;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&accum_full[s], 1);
            mbar_init(&accum_empty[s], 256);
        }
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA_hi);
        prefetch_tmap(&tmB_hi);
        if (NPROD == 3) {
            prefetch_tmap(&tmA_lo);
            prefetch_tmap(&tmB_lo);
        }
    }
    if (warp == 1) {   // TMEM allocation (whole warp): two accumulators of BN fp32 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(2 * BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();                  // programmatic dependent launch: the predecessor's writes are visible from here on (common.cuh)
    pdl_launch_dependents();     // the successor may start its prologue now; it waits the same way before touching memory

    // item -> (k slice, column tile, sample, depth slice, h block, w block); consecutive items share the weight tile
    auto decode = [&](int item, int& kb_begin, int& kb_end, int& n0, int& n, int& od, int& h0, int& w0) {
        // n_fastest (wide outputs, e.g. the [S, S] matrices of the attention backward): CTAs that run at the same time write the
        // column tiles of the SAME rows, so a row's bytes reach L2 / DRAM together instead of 256 B at a time across 148 row tiles
        const int mn = p.m_tiles * p.n_tiles;
        const int in_slice = item % mn;
        const int z = item / mn;
        int t = p.n_fastest ? in_slice / p.n_tiles : in_slice % p.m_tiles;
        n0 = (p.n_fastest ? in_slice % p.n_tiles : in_slice / p.m_tiles) * BN;
        kb_begin = z * p.kb_per_split;
        kb_end = min(total_kb, kb_begin + p.kb_per_split);
        w0 = (t % p.tiles_w) * p.bw; t /= p.tiles_w;
        h0 = (t % p.tiles_h) * p.bh; t /= p.tiles_h;
        od = t % p.d_out;
        n = t / p.d_out;
    };

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            int it = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
                int kb_begin, kb_end, n0, n, od, h0, w0;
                decode(item, kb_begin, kb_end, n0, n, od, h0, w0);
                // k-block -> (tap, channel block) -> (kd, kh, kw): decoded once per item, then walked with carries (the divisions cost more
                // than a single-product stage's worth of tensor cycles)
                int tap = kb_begin / p.cin_blocks, cb = kb_begin - tap * p.cin_blocks;
                int tkw = tap % p.kw, tkh = (tap / p.kw) % p.kh, tkd = tap / (p.kw * p.kh);
                for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
                    const int s = it % Cfg::kStages;
                    const uint32_t ph = (uint32_t)(it / Cfg::kStages) & 1u;
                    mbar_wait(&empty[s], ph ^ 1u);
                    uint8_t* st = smem + s * Cfg::kStageBytes;
                    mbar_expect_tx(&full[s], Cfg::kStageBytes);
                    const int ac = p.a_ch_off + cb * BK, aw = w0 + tkw - p.pw, ah = h0 + tkh - p.ph, ad = od + tkd - p.pd;
                    const int b2 = p.w_batched ? n : tap;
                    tma_load_5d(st, &tmA_hi, &full[s], ac, aw, ah, ad, n);
                    tma_load_3d(st + Cfg::kPlanes * Cfg::kABytes, &tmB_hi, &full[s], cb * BK + p.w_k_off, n0, b2);
                    if (NPROD == 3) {
                        tma_load_5d(st + Cfg::kABytes, &tmA_lo, &full[s], ac, aw, ah, ad, n);
                        tma_load_3d(st + 2 * Cfg::kABytes + Cfg::kBBytes, &tmB_lo, &full[s], cb * BK + p.w_k_off, n0, b2);
                    }
                    if (++cb == p.cin_blocks) {
                        cb = 0;
                        ++tap;
                        if (++tkw == p.kw) {
                            tkw = 0;
                            if (++tkh == p.kh) { tkh = 0; ++tkd; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (one thread) =================
        if (elect_one()) {
            // instruction descriptor: D=f32, A=B=bf16, both K-major, N, M=128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            int it = 0, li = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++li) {
                int kb_begin, kb_end, n0, n, od, h0, w0;
                decode(item, kb_begin, kb_end, n0, n, od, h0, w0);
                const int as = li & 1;
                mbar_wait(&accum_empty[as], (uint32_t)(((li >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(as * BN);
                for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
                    const int first = (kb == kb_begin);
                    const int s = it % Cfg::kStages;
                    const uint32_t ph = (uint32_t)(it / Cfg::kStages) & 1u;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + s * Cfg::kStageBytes);
                    const uint64_t da_hi = make_smem_desc(st);
                    const uint64_t db_hi = make_smem_desc(st + Cfg::kPlanes * Cfg::kABytes);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t koff = (uint64_t)(k * 2);   // 32 bytes >> 4
                        const uint32_t acc_flag = (first && k == 0) ? 0u : 1u;
                        if (NPROD == 3) {
                            const uint64_t da_lo = make_smem_desc(st + Cfg::kABytes);
                            const uint64_t db_lo = make_smem_desc(st + 2 * Cfg::kABytes + Cfg::kBBytes);
                            umma_bf16(tacc, da_lo + koff, db_hi + koff, idesc, acc_flag);
                            umma_bf16(tacc, da_hi + koff, db_lo + koff, idesc, 1u);
                            umma_bf16(tacc, da_hi + koff, db_hi + koff, idesc, 1u);
                        } else {
                            umma_bf16(tacc, da_hi + koff, db_hi + koff, idesc, acc_flag);
                        }
                    }
                    tc_commit(&empty[s]);   // frees the smem stage once these MMAs have read it
                }
                tc_commit(&accum_full[as]);      // accumulator complete
            }
        }
    } else {
        // ================= epilogue warps (2..9): TMEM lanes 32*(warp%4) .. +31, column half (warp-2)/4 =================
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int c_begin = half * (BN / 64), c_end = c_begin + BN / 64;       // this warp's 32-column chunks
        const int row = q * 32 + lane;
        int li = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++li) {
            int kb_begin, kb_end, n0, n, od, h0, w0;
            decode(item, kb_begin, kb_end, n0, n, od, h0, w0);
            const int as = li & 1;
            const int ow = w0 + row % p.bw, oh = h0 + row / p.bw;
            const size_t pos = (((size_t)n * p.d_out + od) * p.h + oh) * p.w + ow;
            const float rv = p.row_mode ? __ldg(p.row_vec + pos) : 0.f;        // in flight while the accumulator completes
            mbar_wait(&accum_full[as], (uint32_t)((li >> 1) & 1));
            tc_fence_after();
            if (TSTORE) {
                // Row tiles (128 consecutive positions): stage the hi / lo tile in swizzled shared memory, one thread hands it to the TMA
                // engine.  A thread-per-row global store touches 32 half-filled sectors per instruction; this path touches none.
                const bool leader = (warp == 2 && lane == 0);
                if (li > 0) {
                    if (leader) tma_store_wait_read();               // the previous tile's stores have read the staging buffer
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
                const uint32_t xr = (uint32_t)(row & 7);
                // all of this thread's accumulator columns leave TMEM first, so the MMA warp gets the buffer back before the epilogue math
                uint32_t accs[BN / 64][32];
#pragma unroll
                for (int k = 0; k < BN / 64; ++k)
                    tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + (c_begin + k) * 32), accs[k]);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&accum_empty[as]);
#pragma unroll
                for (int k = 0; k < BN / 64; ++k) {
                    const int c = c_begin + k;
                    const uint32_t (&acc)[32] = accs[k];
                    float v[32];
                    conv_epilogue_values(p, acc, pos, n0 + c * 32, v, rv);
                    if (p.stats) stat_add_global(p, v, n0 + c * 32, lane);
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
                    const uint32_t blk = (uint32_t)(c >> 1);
                    const uint32_t base_hi = smem_u32(stg) + blk * (BM * 128) + (uint32_t)row * 128;
                    const uint32_t base_lo = base_hi + (BN / 64) * (BM * 128);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const uint32_t off = ((((uint32_t)(c & 1) * 4 + g) ^ xr) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base_hi + off), "r"(hi[4 * g]), "r"(hi[4 * g + 1]),
                                     "r"(hi[4 * g + 2]), "r"(hi[4 * g + 3]) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base_lo + off), "r"(lo[4 * g]), "r"(lo[4 * g + 1]),
                                     "r"(lo[4 * g + 2]), "r"(lo[4 * g + 3]) : "memory");
                    }
                }
                fence_proxy_async();
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (leader) {
                    const int pos0 = (int)(pos - (size_t)row);
#pragma unroll
                    for (int b = 0; b < BN / 64; ++b) {
                        tma_store_2d(&tmO_hi, stg + b * (BM * 128), p.o_ch_off + n0 + 64 * b, pos0);
                        tma_store_2d(&tmO_lo, stg + (BN / 64 + b) * (BM * 128), p.o_ch_off + n0 + 64 * b, pos0);
                    }
                    tma_store_commit();
                }
                continue;
            }
            if (p.coop) {
                // Cooperative split-K (small grids: batch-1 inference, the PRGCN GEMMs): every slice parks its partial tile in the
                // workspace; the warp that arrives LAST at a (tile, row quarter) adds the slices in slice order — a fixed order, so the
                // result does not depend on the arrival order — and runs the normal epilogue.  Counters return to zero.
                const int tile = item % (p.m_tiles * p.n_tiles);
                const int z = item / (p.m_tiles * p.n_tiles);
                const size_t slice_stride = (size_t)p.m_tiles * p.n_tiles * BM * BN;
                float* mine = p.coop_ws + (size_t)z * slice_stride + ((size_t)tile * BM + row) * BN;
#pragma unroll 1
                for (int c = c_begin; c < c_end; ++c) {
                    uint32_t acc[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c * 32), acc);
                    if (c == c_end - 1) {
                        tc_fence_before();
                        mbar_arrive(&accum_empty[as]);
                    }
                    float4* dst = reinterpret_cast<float4*>(mine + c * 32);
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        __stcg(dst + g, make_float4(__uint_as_float(acc[4 * g]), __uint_as_float(acc[4 * g + 1]), __uint_as_float(acc[4 * g + 2]),
                                                    __uint_as_float(acc[4 * g + 3])));
                }
                __threadfence();
                __syncwarp();
                int arrived = 0;
                if (lane == 0) arrived = atomicAdd(p.coop_counters + (tile * 4 + q) * 2 + half, 1);
                arrived = __shfl_sync(0xffffffffu, arrived, 0);
                if (arrived == p.k_split - 1) {
                    __threadfence();
                    const float* first = p.coop_ws + ((size_t)tile * BM + row) * BN;
#pragma unroll 1
                    for (int c = c_begin; c < c_end; ++c) {
                        float sum[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) sum[j] = 0.f;
                        for (int zz = 0; zz < p.k_split; zz += 2) {      // two slices in flight, added in slice order
                            const float4* s0 = reinterpret_cast<const float4*>(first + (size_t)zz * slice_stride + c * 32);
                            const bool two = zz + 1 < p.k_split;
                            const float4* s1 = two ? reinterpret_cast<const float4*>(first + (size_t)(zz + 1) * slice_stride + c * 32) : s0;
                            float4 u[8], w[8];
#pragma unroll
                            for (int g = 0; g < 8; ++g) { u[g] = __ldcg(s0 + g); w[g] = __ldcg(s1 + g); }
#pragma unroll
                            for (int g = 0; g < 8; ++g) {
                                sum[4 * g] += u[g].x; sum[4 * g + 1] += u[g].y; sum[4 * g + 2] += u[g].z; sum[4 * g + 3] += u[g].w;
                            }
                            if (two) {
#pragma unroll
                                for (int g = 0; g < 8; ++g) {
                                    sum[4 * g] += w[g].x; sum[4 * g + 1] += w[g].y; sum[4 * g + 2] += w[g].z; sum[4 * g + 3] += w[g].w;
                                }
                            }
                        }
                        uint32_t acc[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(sum[j]);
                        conv_epilogue32_stats(p, acc, pos, n0 + c * 32, rv, lane);
                    }
                    if (lane == 0) p.coop_counters[(tile * 4 + q) * 2 + half] = 0;
                }
                continue;
            }
            uint32_t accs[BN / 64][32];         // drain this thread's columns, hand the accumulator back, then do the epilogue math
#pragma unroll
            for (int k = 0; k < BN / 64; ++k)
                tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + (c_begin + k) * 32), accs[k]);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&accum_empty[as]);
#pragma unroll
            for (int k = 0; k < BN / 64; ++k) conv_epilogue32_stats(p, accs[k], pos, n0 + (c_begin + k) * 32, rv, lane);
        }
        if (TSTORE && warp == 2 && lane == 0) tma_store_wait_all();      // the CTA's last stores are complete before its smem goes away
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN));
    }
}

// ---- host side -----------------------------------------------------------------------------------
static int encode_act_map(CUtensorMap* map, const void* base, int ca, int w, int h, int d, int n, int bw, int bh, long long n_stride) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return HUPR_ERR_CUDA;
    cuuint64_t dims[5] = {(cuuint64_t)ca, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)d, (cuuint64_t)n};
    cuuint64_t strides[4] = {(cuuint64_t)ca * 2, (cuuint64_t)w * ca * 2, (cuuint64_t)h * w * ca * 2,
                             (cuuint64_t)(n_stride > 0 ? n_stride : (long long)d * h * w * ca) * 2};
    cuuint32_t box[5] = {(cuuint32_t)BK, (cuuint32_t)bw, (cuuint32_t)bh, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? HUPR_OK : HUPR_ERR_CUDA;
}

static int encode_wgt_map(CUtensorMap* map, const void* base, int cin, int cout, int taps, int bn, int ld) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return HUPR_ERR_CUDA;
    cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)cout, (cuuint64_t)taps};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)cout * ld * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)bn, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? HUPR_OK : HUPR_ERR_CUDA;
}

static int encode_out_map(CUtensorMap* map, const void* base, int ld, long long positions) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return HUPR_ERR_CUDA;
    cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)positions};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? HUPR_OK : HUPR_ERR_CUDA;
}

template <int BN, int NPROD, bool TSTORE>
static int launch_conv(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
                       const ConvParams& p, int m_tiles, cudaStream_t stream) {
    using Cfg = ConvCfg<BN, NPROD, TSTORE>;
    static bool configured[kMaxDevices] = {};
    if (int crc = ensure_smem_optin(conv_gemm_kernel<BN, NPROD, TSTORE>, Cfg::kSmemBytes, configured)) return crc;
    const int num_sms = device_sm_count();
    if (num_sms <= 0) return HUPR_ERR_CUDA;
    ConvParams pp = p;
    pp.m_tiles = m_tiles;
    pp.n_tiles = p.cout / BN;
    pp.n_fastest = (pp.n_tiles >= 8 && !getenv("HUPR_M_FASTEST")) ? 1 : 0;
    pp.coop = 0;
    CUtensorMap o_hi = a_hi, o_lo = a_hi;                // placeholders unless TSTORE
    if (TSTORE) {
        const long long positions = (long long)p.n * p.d_out * p.h * p.w;
        int rc;
        if ((rc = encode_out_map(&o_hi, p.o_hi, p.o_ld, positions)) != HUPR_OK) return rc;
        if ((rc = encode_out_map(&o_lo, p.o_lo, p.o_ld, positions)) != HUPR_OK) return rc;
    }
    if (!TSTORE && p.coop_ws && pp.k_split == 1) {     // small grid, long contraction: cooperative split-K with an ordered in-kernel reduction
        const long long tiles = (long long)pp.m_tiles * pp.n_tiles;
        const int total_kb = p.kd * p.kh * p.kw * p.cin_blocks;
        if (tiles * 2 <= num_sms && total_kb >= 8) {
            // up to 8 slices, as many as fill the GPU once; a slice keeps >= 4 k-blocks (>= 2 when the grid is tiny: batch-1 level-3
            // layers and the PRGCN GEMMs stream 4-28 MB of weights through 4-16 tiles, where the serial k-loop of a tile — not the
            // 64 KB park / re-read of its partial sums — is the latency; batch-1 launch list: profiles/r02_b1_launches.csv)
            int ks = (int)(num_sms / tiles);
            const int min_kb = tiles <= 16 ? 2 : 4;
            if (ks > total_kb / min_kb) ks = total_kb / min_kb;
            if (ks > 8) ks = 8;
            const size_t per_slice = (size_t)tiles * BM * BN * sizeof(float);
            const size_t avail = p.coop_ws_bytes > 4096 ? p.coop_ws_bytes - 4096 : 0;
            if ((size_t)ks * per_slice > avail) ks = (int)(avail / per_slice);
            if (ks >= 2 && tiles * 8 * sizeof(int) <= 4096) {
                pp.kb_per_split = (total_kb + ks - 1) / ks;
                pp.k_split = (total_kb + pp.kb_per_split - 1) / pp.kb_per_split;
                pp.coop = pp.k_split > 1 ? 1 : 0;
                pp.coop_counters = reinterpret_cast<int*>(p.coop_ws);
                pp.coop_ws = p.coop_ws + 1024;            // slices start 4096 B in
            }
        }
    }
    const long long items = (long long)pp.m_tiles * pp.n_tiles * pp.k_split;
    if (items > 2147483647LL) return HUPR_ERR_BAD_ARG;
    dim3 grid((unsigned)(items < num_sms ? items : num_sms), 1, 1);
    launch_k(conv_gemm_kernel<BN, NPROD, TSTORE>, grid, dim3(kConvThreads), Cfg::kSmemBytes, stream, a_hi, a_lo, b_hi, b_lo, o_hi, o_lo, pp);
    note_launches(1);
    return cudaGetLastError() == cudaSuccess ? HUPR_OK : HUPR_ERR_CUDA;
}

int conv_halo_try(const hupr_conv_desc* d, const ConvParams& base, cudaStream_t stream, bool probe);   // conv_halo.cu

// Validation + parameter block shared by hupr_conv_gemm and hupr_conv_quant_eligible; `launch` = false stops before any launch and returns
// 1 / 0 for "the two-unit halo kernel would take this descriptor".
static int conv_gemm_impl(const hupr_conv_desc* d, void* stream, bool launch);

}  // namespace hupr

extern "C" int hupr_conv_gemm(const hupr_conv_desc* d, void* stream) { return hupr::conv_gemm_impl(d, stream, true); }

extern "C" int hupr_conv_quant_eligible(const hupr_conv_desc* d) { return hupr::conv_gemm_impl(d, nullptr, false); }

int hupr::conv_gemm_impl(const hupr_conv_desc* d, void* stream, bool launch) {
    using namespace hupr;
    if (!d || !d->a_hi || !d->w_hi) return HUPR_ERR_BAD_ARG;
    if (d->nprod < 0 || d->nprod > 3) return HUPR_ERR_BAD_ARG;
    if (d->nprod != 1 && (d->a_lo == nullptr) != (d->w_lo == nullptr)) return HUPR_ERR_BAD_ARG;
    if (d->nprod >= 2 && !d->a_lo) return HUPR_ERR_BAD_ARG;
    {   // the quantised planes come in complete sets
        const int na = (d->a_q16 != nullptr) + (d->a_q8 != nullptr);
        const int nw = (d->w_q16 != nullptr) + (d->w_q8 != nullptr);
        const int no = (d->o_q16 != nullptr) + (d->o_q8 != nullptr);
        if (na % 2 || nw % 2 || no % 2 || (no && !d->o_hi)) return HUPR_ERR_BAD_ARG;
        if (no && (d->o_ld % 32 || d->o_ch_off % 32)) return HUPR_ERR_BAD_ARG;      // 32-channel blocks of the e4m3 plane
        if (d->nprod == 2 && launch && (!na || !nw)) return HUPR_ERR_BAD_ARG;
        if (no && (d->k_split > 1)) return HUPR_ERR_BAD_ARG;
    }
    if (d->n <= 0 || d->d <= 0 || d->h <= 0 || d->w <= 0) return HUPR_ERR_BAD_ARG;
    if (d->a_n_stride < 0 || d->a_n_stride % 8) return HUPR_ERR_BAD_ARG;
    if (d->ca % 8 || d->cin % BK || d->cin <= 0 || d->a_ch_off % 8 || d->a_ch_off + d->cin > ((d->ca + BK - 1) / BK) * BK) return HUPR_ERR_BAD_ARG;
    if (d->cout <= 0 || d->cout % 64) return HUPR_ERR_BAD_ARG;
    if (d->kd <= 0 || d->kh <= 0 || d->kw <= 0) return HUPR_ERR_BAD_ARG;
    if (d->kh != 2 * d->ph + 1 || d->kw != 2 * d->pw + 1) return HUPR_ERR_BAD_ARG;   // H, W are 'same' convolutions
    const int d_out = d->d + 2 * d->pd - d->kd + 1;
    if (d_out <= 0) return HUPR_ERR_BAD_ARG;
    const int taps = d->kd * d->kh * d->kw;
    if (d->w_batched && taps != 1) return HUPR_ERR_BAD_ARG;
    if (!d->o_hi && !d->o_f32) return HUPR_ERR_BAD_ARG;
    if (d->o_hi && (d->o_ld % 8 || d->o_ch_off % 8 || d->o_ch_off + d->cout > d->o_ld)) return HUPR_ERR_BAD_ARG;
    if (d->o_f32 && (d->o_f32_ld % 4 || d->cout > d->o_f32_ld)) return HUPR_ERR_BAD_ARG;
    if (d->r_hi && (d->r_ld % 8 || d->r_ch_off % 8 || d->r_ch_off + d->cout > d->r_ld)) return HUPR_ERR_BAD_ARG;
    // spatial tiling of the 128 output positions of one CTA
    const int bw = d->w < BM ? d->w : BM;
    if (BM % bw || d->w % bw) return HUPR_ERR_BAD_ARG;
    const int bh = BM / bw;
    if (d->h % bh) return HUPR_ERR_BAD_ARG;
    const uintptr_t align_or = (uintptr_t)d->a_hi | (uintptr_t)d->a_lo | (uintptr_t)d->w_hi | (uintptr_t)d->w_lo | (uintptr_t)d->o_hi |
                               (uintptr_t)d->o_lo | (uintptr_t)d->o_f32 | (uintptr_t)d->r_hi | (uintptr_t)d->r_lo | (uintptr_t)d->a_q16 |
                               (uintptr_t)d->a_q8 | (uintptr_t)d->w_q16 | (uintptr_t)d->w_q8 | (uintptr_t)d->o_q16 | (uintptr_t)d->o_q8;
    if (align_or & 15) return HUPR_ERR_ALIGNMENT;

    ConvParams p;
    p.n = d->n; p.d_in = d->d; p.h = d->h; p.w = d->w; p.d_out = d_out;
    p.kd = d->kd; p.kh = d->kh; p.kw = d->kw; p.pd = d->pd; p.ph = d->ph; p.pw = d->pw;
    p.cin_blocks = d->cin / BK; p.a_ch_off = d->a_ch_off; p.w_batched = d->w_batched;
    p.bw = bw; p.bh = bh; p.tiles_w = d->w / bw; p.tiles_h = d->h / bh;
    p.cout = d->cout;
    p.scale = d->scale; p.shift = d->shift; p.slope = d->slope;
    p.r_hi = (const __nv_bfloat16*)d->r_hi; p.r_lo = (const __nv_bfloat16*)d->r_lo; p.r_ld = d->r_ld; p.r_ch_off = d->r_ch_off;
    p.o_hi = (__nv_bfloat16*)d->o_hi; p.o_lo = (__nv_bfloat16*)d->o_lo; p.o_ld = d->o_ld; p.o_ch_off = d->o_ch_off;
    p.o_f32 = d->o_f32; p.o_f32_ld = d->o_f32_ld;
    p.w_k_off = d->w_k_off;
    p.row_vec = d->row_vec; p.row_mode = d->row_vec ? d->row_mode : 0;
    p.stats = d->stats; p.stats_ld = d->stats_ld;
    p.acc_scale = 1.0f;
    p.o_q16 = static_cast<__half*>(d->o_q16); p.o_q8 = static_cast<uint8_t*>(d->o_q8);
    p.q_sat = d->q_sat;
    {   // 256-bit stores need 32-byte aligned row segments in every plane that is written (e4m3 planes: one byte per element)
        const uintptr_t o_or = (uintptr_t)d->o_hi | (uintptr_t)d->o_lo | (uintptr_t)d->o_q16 | (uintptr_t)d->o_q8;
        const int unit = d->o_q8 ? 32 : 16;
        p.st256 = (d->o_hi && !(o_or & 31) && d->o_ld % unit == 0 && d->o_ch_off % unit == 0 && !getenv("HUPR_NO_ST256")) ? 1 : 0;
    }
    if (d->stats && (d->stats_ld < d->cout || d->k_split > 1 || ((uintptr_t)d->stats & 7))) return HUPR_ERR_BAD_ARG;
    p.coop_ws = static_cast<float*>(d->ws); p.coop_ws_bytes = d->ws ? d->ws_bytes : 0; p.coop_counters = nullptr; p.coop = 0;
    if (d->ws && ((uintptr_t)d->ws & 15)) return HUPR_ERR_ALIGNMENT;
    if (p.row_mode < 0 || p.row_mode > 2 || (p.row_mode && (d->scale || d->shift || d->slope || !d->o_hi)) || (p.row_mode == 2 && !d->r_hi)) return HUPR_ERR_BAD_ARG;
    {   // split-K (wgrad-style contractions: few output tiles, very long K): partial sums are added atomically into o_f32
        const int total_kb = taps * p.cin_blocks;
        int ks = d->k_split > 1 ? d->k_split : 1;
        if (ks > total_kb) ks = total_kb;
        if (ks > 1 && (!d->o_f32 || d->o_hi || d->scale || d->shift || d->slope || d->r_hi)) return HUPR_ERR_BAD_ARG;
        p.kb_per_split = (total_kb + ks - 1) / ks;
        p.k_split = (total_kb + p.kb_per_split - 1) / p.kb_per_split;
        p.atomic = p.k_split > 1 ? 1 : 0;
    }

    {   // 3-tap-in-H convolutions with enough tiles go to the halo-reuse kernel (conv_halo.cu)
        const int hr = conv_halo_try(d, p, static_cast<cudaStream_t>(stream), !launch);
        if (!launch) return hr == 2 ? 1 : (hr < 0 ? hr : 0);      // probe: 2 = the two-unit kernel takes it
        if (hr <= 0) return hr;
    }
    const int bn = (d->cout % 128 == 0) ? 128 : 64;
    const int wdim2 = d->w_batched ? d->n : taps;
    CUtensorMap a_hi, a_lo, b_hi, b_lo;
    int rc;
    if ((rc = encode_act_map(&a_hi, d->a_hi, d->ca, d->w, d->h, d->d, d->n, bw, bh, d->a_n_stride)) != HUPR_OK) return rc;
    const int w_ld = d->w_ld ? d->w_ld : d->cin;
    if (w_ld % 8 || d->w_ch_off % 8 || d->w_ch_off < 0 || d->w_ch_off + d->cin > w_ld) return HUPR_ERR_BAD_ARG;
    const __nv_bfloat16* w_hi = static_cast<const __nv_bfloat16*>(d->w_hi) + d->w_ch_off;
    const __nv_bfloat16* w_lo = d->w_lo ? static_cast<const __nv_bfloat16*>(d->w_lo) + d->w_ch_off : nullptr;
    if ((rc = encode_wgt_map(&b_hi, w_hi, d->cin, d->cout, wdim2, bn, w_ld)) != HUPR_OK) return rc;
    const bool split = d->a_lo != nullptr && d->w_lo != nullptr && d->nprod != 1;      // nprod == 1: the lo planes are not read
    if (split) {
        if ((rc = encode_act_map(&a_lo, d->a_lo, d->ca, d->w, d->h, d->d, d->n, bw, bh, d->a_n_stride)) != HUPR_OK) return rc;
        if ((rc = encode_wgt_map(&b_lo, w_lo, d->cin, d->cout, wdim2, bn, w_ld)) != HUPR_OK) return rc;
    } else {
        a_lo = a_hi;
        b_lo = b_hi;
    }
    const long long m_tiles_ll = (long long)d->n * d_out * p.tiles_h * p.tiles_w;
    if (m_tiles_ll > 2147483647LL) return HUPR_ERR_BAD_ARG;
    const int m_tiles = (int)m_tiles_ll;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // row tiles with bf16 split output and enough tiles to fill the GPU: TMA-store epilogue
    const bool tstore = bw == BM && d->o_hi && d->o_lo && !d->o_f32 && !p.atomic && m_tiles_ll * (d->cout / bn) >= 148 &&
                        (long long)d->n * d_out * d->h * d->w < 2147483647LL && !d->no_tma_store && !d->o_q16;
    if (tstore && split) return bn == 128 ? launch_conv<128, 3, true>(a_hi, a_lo, b_hi, b_lo, p, m_tiles, s) : launch_conv<64, 3, true>(a_hi, a_lo, b_hi, b_lo, p, m_tiles, s);
    if (tstore) return bn == 128 ? launch_conv<128, 1, true>(a_hi, a_lo, b_hi, b_lo, p, m_tiles, s) : launch_conv<64, 1, true>(a_hi, a_lo, b_hi, b_lo, p, m_tiles, s);
    if (bn == 128) return split ? launch_conv<128, 3, false>(a_hi, a_lo, b_hi, b_lo, p, m_tiles, s) : launch_conv<128, 1, false>(a_hi, a_lo, b_hi, b_lo, p, m_tiles, s);
    return split ? launch_conv<64, 3, false>(a_hi, a_lo, b_hi, b_lo, p, m_tiles, s) : launch_conv<64, 1, false>(a_hi, a_lo, b_hi, b_lo, p, m_tiles, s);
}
