// 3-tap-in-H implicit-GEMM convolution with halo reuse (tcgen05 + TMEM + TMA), sm_100a.
//
// Same contract as conv_gemm.cu (hupr_conv_gemm dispatches here for kh == 3 split-bf16 convolutions, i.e. every 3x3x3
// Conv3d and 3x3 Conv2d of MSCSA-PRGCN, /root/reference/models/layers.py:24-32,45-63).  The generic kernel re-reads the
// activation tile once per filter tap and the weight tile once per 128 output rows.  This kernel:
//   * one CTA owns 256 output positions (bh x bw, two 128-row TMEM accumulators) so every weight tile feeds two MMAs;
//   * for each (kd, kw, 32-channel block) ONE activation box of bh+2 rows is loaded; the three kh taps of both
//     accumulators are sub-views of it (start address + kh*bw rows — legal because bw*64 B is a multiple of the swizzle
//     atom), i.e. 6 tile-taps per load instead of 1;
//   * k-blocks are 32 channels (64-byte rows, SWIZZLE_64B) so that a stage (halo box + 3 weight tiles, two planes) stays under
//     96 KiB and two to four stages fit; one TMA box brings the three kh taps of a weight plane and, when lo lies above hi in the
//     address space (SplitTensor: always), both planes (weight map dims cin, cout, plane, kw, kd*3 + kh);
//   * the kernel is persistent (one CTA per SM walks the tiles) with two accumulator sets in TMEM, so the epilogue of one tile
//     (TMEM -> scale/shift/residual/activation -> hi/lo stores) overlaps the tensor-core main loop of the next; EIGHT epilogue
//     warps, four per accumulator of the tile, and 256-bit stores (conv_common.cuh): a thread owns one output row, and 16-byte
//     stores wrote every 32-byte sector in two instructions — the first two-unit kernel spent 77 k cycles per tile in the epilogue
//     against a 37 k-cycle main loop (ncu: profiles/r02_halo128_q_*);
//   * what bounds the main loop is SHARED-MEMORY BANDWIDTH: an M128 x N128 x K16 MMA with both operands in shared memory reads
//     4 KB + 4 KB in its 64 cycles = the whole 128 B/clk port, and the TMA fills of the next stage come on top.
//     TWO = true runs CTA PAIRS (tcgen05.mma.cta_group::2, tc.cuh): the two CTAs of a cluster take neighbouring position tiles of the
//     same weight column tile, each keeps HALF of the weight rows, the leader issues M = 256 MMAs that fill both CTAs'
//     accumulators — 6 KB per MMA and 72 KB per stage and CTA: tensor pipe 76.8 -> 92.7 % active (three products);
//   * cout = 64 tiles (BN = 64, three products): an M128 x N64 MMA reads 4 KB of A for 32 tensor cycles.  The weight tiles of a
//     tap are used as one 128-row B operand [w_hi | w_lo], so a_hi meets both in ONE N = 128 MMA (columns 0..63 = hi*hi,
//     64..127 = hi*lo) and a_lo * w_hi is a second, N = 64 MMA into columns 0..63; the epilogue adds the two column halves.
//     In pairs the leader holds the w_hi tile and the peer the w_lo tile of that operand, plus a 32-row half of w_hi each;
//   * NPROD == 2 is the two-unit arithmetic of hupr_conv_desc (include/hupr_b200.h): per 32-channel block ONE kind::f16 pair of
//     MMAs on fp16 planes (a16 x w16) plus TWO kind::f8f6f4 MMAs (K = 32) on e4m3 operands (al x w, a x wl), all into the same
//     fp32 accumulator at a common 2^16 scale — 2/3 of the tensor cycles of three bf16 products.  The two e4m3 planes of a tensor
//     are interleaved per 32-channel block ([values | residuals] = 64-byte rows, SWIZZLE_64B like the fp16 plane; as separate
//     32-byte-row planes under SWIZZLE_32B the tensor core read them at half the rate): the two cross-term operands are the K = 32
//     slices at byte 0 and byte 32 of the same rows.  A stage has the size of the hi/lo one.
// Warp roles and the hi/lo 3-product arithmetic are those of conv_gemm.cu.
#include <stdlib.h>

#include "tc.cuh"
#include "conv_common.cuh"

namespace hupr {

constexpr int HK = 32;             // channels per k-block (64-byte swizzled rows)
constexpr int HM = 256;            // output positions per CTA
constexpr int kHaloThreads = 320;                    // TMA warp, MMA warp, 2 x 4 epilogue warps (one group per accumulator of a tile)
constexpr int kHaloStatBytes = 8 * 2 * 128 * 4;      // per-warp running column sums of the fused BatchNorm statistics

struct HaloGeom {
    int halo_rows;                 // (bh + 2) * bw
    int a_plane_bytes;             // halo_rows * 64
    int b_tile_bytes;              // weight rows held by one CTA (BN, or BN / 2 in a CTA pair) * 64
    int nprod;                     // 3: hi/lo planes, three products per k-step; 1: hi planes only
    int stage_bytes;               // planes * (a_plane_bytes + 3 * b_tile_bytes), planes = 2 (nprod 3) or 1
    int stages;
    int tap_stride_bytes;          // bw * 64: one h-row of the halo
    int acc_stride_bytes;          // (bh / 2) * bw * 64: first row of accumulator 1
    int m_tiles, n_tiles;          // 256-position tiles x BN-column tiles, walked persistently
    int b_merged;                  // three products: tmB_hi is a 5-D map over BOTH weight planes (cin, cout, plane, kw, kd*3 + kh): one box per stage
};

__device__ __forceinline__ uint64_t make_smem_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(512 >> 4) << 32;                   // stride byte offset: 8 rows * 64 B
    d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
    d |= (uint64_t)4 << 61;                            // SWIZZLE_64B
    return d;
}

template <int BN, int NPROD, bool TWO>
__global__ void __launch_bounds__(kHaloThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmA_x,      // NPROD == 2: A maps are (fp16 a16, interleaved e4m3 plane, -), B maps likewise;
                                                                 // cout = 64 pairs: tmB_x = the 32-row half of w_hi
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                 const __grid_constant__ CUtensorMap tmB_x, const ConvParams p, const HaloGeom g) {
    // Persistent: CTA b walks tiles b, b + gridDim.x, ...; the TMA->MMA smem ring runs continuously across tiles and the two
    // accumulator SETS in TMEM (2 x [2 x BN] columns) let the epilogue of tile i overlap the main loop of tile i + 1.
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.stages * g.stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + 4;
    uint64_t* accum_full = bars + 8;     // [2]
    uint64_t* accum_empty = bars + 10;   // [2]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 12);
    float* s_stat = reinterpret_cast<float*>(bars + 32);      // [8 epilogue warps][2][BN] running column sums (fused BatchNorm statistics)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int groups = p.kd * p.kw * p.cin_blocks;      // one stage per (kd, kw, channel block): 3 kh taps x 2 accumulators
    constexpr bool three = NPROD == 3;      // compile-time: the single-thread MMA issue loop carries no runtime branches
    constexpr bool quant = NPROD == 2;      // fp16 main product + two e4m3 cross products (header comment)
    // TWO: CTA pairs (tc.cuh).  The pair works on two neighbouring position tiles (tile = blockIdx.x + i * gridDim.x, cluster ranks = blockIdx.x
    // parity) that share their weight column tile; each CTA keeps the BN / 2 weight rows n0 + rank * BN / 2 ... and the leader's M = 256 MMAs
    // fill both CTAs' accumulators.
    const uint32_t cta_rank = TWO ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    constexpr int BROWS = TWO ? BN / 2 : BN;
    static_assert(!TWO || NPROD != 2 || BN == 128, "two-unit arithmetic: cout = 128 tiles");
    constexpr bool ncat = (BN == 64 && NPROD == 3);         // [w_hi | w_lo] as one N = 128 operand (header comment)
    constexpr int ACC = ncat ? 128 : BN;                    // TMEM columns of one accumulator
    constexpr int TMEM_COLS = 4 * ACC;                      // two sets x two accumulators
    const int total_tiles = g.m_tiles * g.n_tiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < g.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&accum_full[s], 1);
            mbar_init(&accum_empty[s], TWO ? 512 : 256);      // TWO: the leader's barrier collects both CTAs' epilogue threads
        }
        fence_mbar_init();
        prefetch_tmap(&tmA_hi); prefetch_tmap(&tmA_lo); prefetch_tmap(&tmB_hi); prefetch_tmap(&tmB_lo);
        prefetch_tmap(&tmB_x);
    }
    if (warp == 1) {
        if (TWO) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (TWO) cluster_sync_all();      // both CTAs' barriers are initialised and their tensor memory allocated before anything crosses over
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();                  // programmatic dependent launch: the predecessor's writes are visible from here on (common.cuh)
    pdl_launch_dependents();     // the successor may start its prologue now; it waits the same way before touching memory

    // tile -> (weight column tile, sample, depth slice, h block); tiles that run concurrently share the weight tile
    auto decode = [&](int tile, int& n0, int& n, int& od, int& h0) {
        n0 = (tile / g.m_tiles) * BN;
        int t = tile % g.m_tiles;
        h0 = (t % p.tiles_h) * p.bh; t /= p.tiles_h;
        od = t % p.d_out;
        n = t / p.d_out;
    };

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            int it = 0, sidx = 0;
            uint32_t sph = 0;                              // stage ring position / phase, advanced without divisions
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int n0, n, od, h0;
                decode(tile, n0, n, od, h0);
                int cb = 0, tkw = 0, tkd = 0;                 // gi = (tkd * kw + tkw) * cin_blocks + cb, walked with carries (no divisions per stage)
                for (int gi = 0; gi < groups; ++gi, ++it) {
                    const int s = sidx;
                    mbar_wait(&empty[s], sph ^ 1u);
                    uint8_t* st = smem + s * g.stage_bytes;
                    const int ac = p.a_ch_off + cb * HK;
                    if (TWO) {      // both CTAs' loads complete on the leader's barrier, which expects the bytes of both
                        if (leader) mbar_expect_tx(&full[s], 2u * (uint32_t)g.stage_bytes);
                        uint8_t* sb2 = st + (three ? 2 : 1) * g.a_plane_bytes;
                        const int nrow = n0 + (int)cta_rank * BROWS;
                        tma2_load_5d(st, &tmA_hi, &full[s], ac, tkw - p.pw, h0 - 1, od + tkd - p.pd, n);
                        if (ncat) {
                            // cout = 64 in pairs: the N = 128 operand [w_hi | w_lo] splits into the leader's w_hi tile and the peer's w_lo tile
                            // (X, 64 rows each), the N = 64 operand w_hi into two 32-row halves (Y): stage B region = [X0 X1 X2 | Y0 Y1 Y2]
                            tma2_load_5d(st + g.a_plane_bytes, &tmA_lo, &full[s], ac, tkw - p.pw, h0 - 1, od + tkd - p.pd, n);
                            tma2_load_4d(sb2, leader ? &tmB_hi : &tmB_lo, &full[s], cb * HK, n0, tkw, tkd * 3);
                            tma2_load_4d(sb2 + 3 * g.b_tile_bytes, &tmB_x, &full[s], cb * HK, n0 + (int)cta_rank * 32, tkw, tkd * 3);
                        } else if (quant) {
                            tma2_load_5d(st + g.a_plane_bytes, &tmA_lo, &full[s], 2 * ac, tkw - p.pw, h0 - 1, od + tkd - p.pd, n);
                            uint8_t* sq = st + 2 * g.a_plane_bytes;
                            tma2_load_4d(sq, &tmB_hi, &full[s], cb * HK, nrow, tkw, tkd * 3);
                            tma2_load_4d(sq + 3 * g.b_tile_bytes, &tmB_lo, &full[s], 2 * cb * HK, nrow, tkw, tkd * 3);
                        } else if (three) {
                            tma2_load_5d(st + g.a_plane_bytes, &tmA_lo, &full[s], ac, tkw - p.pw, h0 - 1, od + tkd - p.pd, n);
                            if (g.b_merged) {
                                tma2_load_5d(sb2, &tmB_hi, &full[s], cb * HK, nrow, 0, tkw, tkd * 3);
                            } else {
#pragma unroll
                                for (int tkh = 0; tkh < 3; ++tkh) {
                                    tma2_load_4d(sb2 + 2 * tkh * g.b_tile_bytes, &tmB_hi, &full[s], cb * HK, nrow, tkw, tkd * 3 + tkh);
                                    tma2_load_4d(sb2 + (2 * tkh + 1) * g.b_tile_bytes, &tmB_lo, &full[s], cb * HK, nrow, tkw, tkd * 3 + tkh);
                                }
                            }
                        } else {
                            tma2_load_4d(sb2, &tmB_hi, &full[s], cb * HK, nrow, tkw, tkd * 3);
                        }
                        if (++cb == p.cin_blocks) {
                            cb = 0;
                            if (++tkw == p.kw) { tkw = 0; ++tkd; }
                        }
                        if (++sidx == g.stages) { sidx = 0; sph ^= 1u; }
                        continue;
                    }
                    mbar_expect_tx(&full[s], (uint32_t)g.stage_bytes);
                    tma_load_5d(st, &tmA_hi, &full[s], ac, tkw - p.pw, h0 - 1, od + tkd - p.pd, n);
                    if (three) tma_load_5d(st + g.a_plane_bytes, &tmA_lo, &full[s], ac, tkw - p.pw, h0 - 1, od + tkd - p.pd, n);
                    if (quant) {       // stage = [a16 | a8x | w16 x3 | w8x x3]; the e4m3 maps count bytes: 64 per 32-channel block
                        tma_load_5d(st + g.a_plane_bytes, &tmA_lo, &full[s], 2 * ac, tkw - p.pw, h0 - 1, od + tkd - p.pd, n);
                        uint8_t* sq = st + 2 * g.a_plane_bytes;
                        // one box per plane brings the three kh taps (weight map dims: cin, cout, kw, kd*3 + kh)
                        tma_load_4d(sq, &tmB_hi, &full[s], cb * HK, n0, tkw, tkd * 3);
                        tma_load_4d(sq + 3 * g.b_tile_bytes, &tmB_lo, &full[s], 2 * cb * HK, n0, tkw, tkd * 3);
                    }
                    uint8_t* sb = st + (three ? 2 : 1) * g.a_plane_bytes;
                    // The TMA unit ingests about one box row per clock plus ~110 clocks per box (tools_dev/micro/tma_rate.cu).  Weight maps have
                    // (kw, kd*3 + kh) as separate dimensions so that one box brings the three kh taps, and, where the two planes can be addressed
                    // as one tensor, both planes (fewer instructions for the producer thread; the stage time itself is set by shared-memory
                    // bandwidth, not by TMA issue — merging the boxes changed nothing measurable):
                    // B tiles of a three-product stage: [hi0 lo0 hi1 lo1 hi2 lo2] (for cout = 64, hi|lo of a tap form one N = 128 operand)
                    if (three) {
                        if (g.b_merged) {
                            tma_load_5d(sb, &tmB_hi, &full[s], cb * HK, n0, 0, tkw, tkd * 3);
                        } else {
#pragma unroll
                            for (int tkh = 0; tkh < 3; ++tkh) {
                                tma_load_4d(sb + 2 * tkh * g.b_tile_bytes, &tmB_hi, &full[s], cb * HK, n0, tkw, tkd * 3 + tkh);
                                tma_load_4d(sb + (2 * tkh + 1) * g.b_tile_bytes, &tmB_lo, &full[s], cb * HK, n0, tkw, tkd * 3 + tkh);
                            }
                        }
                    } else if (!quant) {
                        tma_load_4d(sb, &tmB_hi, &full[s], cb * HK, n0, tkw, tkd * 3);      // [hi0 hi1 hi2]
                    }
                    if (++cb == p.cin_blocks) {
                        cb = 0;
                        if (++tkw == p.kw) { tkw = 0; ++tkd; }
                    }
                    if (++sidx == g.stages) { sidx = 0; sph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (one thread) =================
        if ((!TWO || leader) && elect_one()) {
            // fp32 accumulate; A / B formats: bf16 (1), or for NPROD == 2 fp16 under kind::f16 and e4m3 under kind::f8f6f4 (both 0)
            const uint32_t fmt = quant ? 0u : ((1u << 7) | (1u << 10));
            const uint32_t idesc = (1u << 4) | fmt | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((TWO ? 256 : 128) >> 4) << 24);
            const uint32_t idesc_cat = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)((TWO ? 256 : 128) >> 4) << 24);      // N = 128
            int it = 0, lt = 0, sidx = 0;
            uint32_t sph = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
                const int as = lt & 1;
                mbar_wait(&accum_empty[as], (uint32_t)(((lt >> 1) & 1) ^ 1));     // the epilogue drained this accumulator set
                tc_fence_after();
                const uint32_t tset = tmem_base + (uint32_t)(as * 2 * ACC);
                for (int gi = 0; gi < groups; ++gi, ++it) {
                    const int s = sidx;
                    mbar_wait(&full[s], sph);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + s * g.stage_bytes);
                    const uint32_t sb = st + (three ? 2 : 1) * g.a_plane_bytes;
                    if (quant) {
                        const uint32_t sq = st + 2 * g.a_plane_bytes;
#pragma unroll
                        for (int tkh = 0; tkh < 3; ++tkh) {
                            const uint64_t db16 = make_smem_desc_sw64(sq + tkh * g.b_tile_bytes);
                            const uint64_t db8 = make_smem_desc_sw64(sq + (3 + tkh) * g.b_tile_bytes);      // values: bytes 0..31 of the 64-byte rows
                            const uint64_t db8l = db8 + 2;                                                   // residuals: bytes 32..63
#pragma unroll
                            for (int a = 0; a < 2; ++a) {
                                const uint32_t aoff = (uint32_t)(tkh * g.tap_stride_bytes + a * g.acc_stride_bytes);
                                const uint64_t da16 = make_smem_desc_sw64(st + aoff);
                                const uint64_t da8 = make_smem_desc_sw64(st + g.a_plane_bytes + aoff);
                                const uint64_t da8l = da8 + 2;
                                const uint32_t tacc = tset + (uint32_t)(a * ACC);
                                if (TWO) {
                                    umma2_f8(tacc, da8l, db8, idesc, (uint32_t)((gi | tkh) != 0));
                                    umma2_f8(tacc, da8, db8l, idesc, 1u);
                                    umma2_bf16(tacc, da16, db16, idesc, 1u);
                                    umma2_bf16(tacc, da16 + 2, db16 + 2, idesc, 1u);
                                } else {
                                    umma_f8(tacc, da8l, db8, idesc, (uint32_t)((gi | tkh) != 0));      // (a - a16) x w   (e4m3, K = 32)
                                    umma_f8(tacc, da8, db8l, idesc, 1u);                                // a x (w - w16)
                                    umma_bf16(tacc, da16, db16, idesc, 1u);                             // a16 x w16 (fp16, K = 16 twice)
                                    umma_bf16(tacc, da16 + 2, db16 + 2, idesc, 1u);
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int tkh = 0; tkh < (quant ? 0 : 3); ++tkh) {
                        // (pairs with cout = 64: db_hi = this tap's X tile, db_lo = its 32-row Y tile — see the producer)
                        const uint64_t db_hi = make_smem_desc_sw64(sb + ((TWO && ncat) ? tkh : three ? 2 * tkh : tkh) * g.b_tile_bytes);
                        const uint64_t db_lo = make_smem_desc_sw64(sb + ((TWO && ncat) ? 3 * g.b_tile_bytes + tkh * (g.b_tile_bytes / 2)
                                                                                      : (2 * tkh + 1) * g.b_tile_bytes));
#pragma unroll
                        for (int a = 0; a < 2; ++a) {
                            const uint32_t aoff = (uint32_t)(tkh * g.tap_stride_bytes + a * g.acc_stride_bytes);
                            const uint64_t da_hi = make_smem_desc_sw64(st + aoff);
                            const uint64_t da_lo = make_smem_desc_sw64(st + g.a_plane_bytes + aoff);
                            const uint32_t tacc = tset + (uint32_t)(a * ACC);
#pragma unroll
                            for (int k = 0; k < HK / 16; ++k) {
                                const uint64_t koff = (uint64_t)(k * 2);   // 32 bytes >> 4
                                const uint32_t acc_flag = (gi | tkh | k) != 0;
                                if (ncat && TWO) {
                                    umma2_bf16(tacc, da_hi + koff, db_hi + koff, idesc_cat, acc_flag);     // a_hi x [w_hi (leader) | w_lo (peer)]
                                    umma2_bf16(tacc, da_lo + koff, db_lo + koff, idesc, 1u);               // a_lo x w_hi (32 rows from each CTA)
                                } else if (ncat) {
                                    umma_bf16(tacc, da_hi + koff, db_hi + koff, idesc_cat, acc_flag);      // a_hi x [w_hi | w_lo] -> columns 0..127
                                    umma_bf16(tacc, da_lo + koff, db_hi + koff, idesc, 1u);                // a_lo x w_hi -> columns 0..63
                                } else if (three) {
                                    if (TWO) {
                                        umma2_bf16(tacc, da_lo + koff, db_hi + koff, idesc, acc_flag);
                                        umma2_bf16(tacc, da_hi + koff, db_lo + koff, idesc, 1u);
                                        umma2_bf16(tacc, da_hi + koff, db_hi + koff, idesc, 1u);
                                    } else {
                                        umma_bf16(tacc, da_lo + koff, db_hi + koff, idesc, acc_flag);
                                        umma_bf16(tacc, da_hi + koff, db_lo + koff, idesc, 1u);
                                        umma_bf16(tacc, da_hi + koff, db_hi + koff, idesc, 1u);
                                    }
                                } else if (TWO) {
                                    umma2_bf16(tacc, da_hi + koff, db_hi + koff, idesc, acc_flag);
                                } else {
                                    umma_bf16(tacc, da_hi + koff, db_hi + koff, idesc, acc_flag);
                                }
                            }
                        }
                    }
                    if (TWO) tc_commit2(&empty[s]); else tc_commit(&empty[s]);
                    if (++sidx == g.stages) { sidx = 0; sph ^= 1u; }
                }
                if (TWO) tc_commit2(&accum_full[as]); else tc_commit(&accum_full[as]);
            }
        }
    } else {
        // ================= epilogue warps: TMEM lanes 32*(warp%4) .. +31; warps 2..5 drain accumulator 0, warps 6..9 accumulator 1 ====
        const int q = warp & 3;
        const int a = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        int lt = 0;
        // fused BatchNorm statistics: this warp's running column sums (sum v, sum v^2 of channel n0 + col over all rows it has seen) live
        // in shared memory, lane j owning columns j, j + 32, ...; a CTA's tiles come in non-decreasing n0 order, so the sums are flushed
        // (double atomics) when n0 changes and at the end
        float* wstat = s_stat + (warp - 2) * 2 * BN;
        if (p.stats) {
            for (int j = lane; j < 2 * BN; j += 32) wstat[j] = 0.f;
        }
        int st_n0 = -1;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
            int n0, n, od, h0;
            decode(tile, n0, n, od, h0);
            if (p.stats && n0 != st_n0) {
                if (st_n0 >= 0) stat_flush(p, wstat, BN, st_n0, lane);
                st_n0 = n0;
            }
            const int as = lt & 1;
            mbar_wait(&accum_full[as], (uint32_t)((lt >> 1) & 1));
            tc_fence_after();
            const uint32_t tset = tmem_base + (uint32_t)(as * 2 * ACC) + ((uint32_t)(q * 32) << 16);
            {
                const int ow = row % p.bw, oh = h0 + a * (p.bh / 2) + row / p.bw;
                const size_t pos = (((size_t)n * p.d_out + od) * p.h + oh) * p.w + ow;
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t acc[32];
                    if (ncat) {                            // hi*hi + lo*hi (columns c) + hi*lo (columns 64 + c)
                        uint32_t cross[32];
                        tmem_ld32_nowait(tset + (uint32_t)(a * ACC + c * 32), acc);
                        tmem_ld32_nowait(tset + (uint32_t)(a * ACC + 64 + c * 32), cross);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(__uint_as_float(acc[j]) + __uint_as_float(cross[j]));
                    } else {
                        tmem_ld32(tset + (uint32_t)(a * BN + c * 32), acc);
                    }
                    if (c == BN / 32 - 1) {                // this thread's last TMEM read of the set: hand it back before the stores
                        tc_fence_before();
                        if (TWO) mbar_arrive_leader(&accum_empty[as]); else mbar_arrive(&accum_empty[as]);
                    }
                    conv_epilogue32(p, acc, pos, n0 + c * 32, p.row_mode ? __ldg(p.row_vec + pos) : 0.f, wstat + c * 32, BN, lane);
                }
            }
        }
        if (p.stats && st_n0 >= 0) stat_flush(p, wstat, BN, st_n0, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (TWO) cluster_sync_all();      // the leader's MMAs read the peer's shared memory and write its tensor memory until its last tile is done
    if (warp == 1) {
        if (TWO) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

// esize: bytes per element — 2: bf16 / fp16 planes, 1: the interleaved e4m3 plane (a byte tensor with 2 x channels per row); 64-byte rows
static int encode_halo_act_map(CUtensorMap* map, const void* base, int ca, int w, int h, int d, int n, int bw, int box_h, long long n_stride,
                               int esize = 2) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return HUPR_ERR_CUDA;
    const cuuint64_t e = (cuuint64_t)esize;
    cuuint64_t dims[5] = {(cuuint64_t)ca, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)d, (cuuint64_t)n};
    cuuint64_t strides[4] = {(cuuint64_t)ca * e, (cuuint64_t)w * ca * e, (cuuint64_t)h * w * ca * e,
                             (cuuint64_t)(n_stride > 0 ? n_stride : (long long)d * h * w * ca) * e};
    cuuint32_t box[5] = {(cuuint32_t)(64 / esize), (cuuint32_t)bw, (cuuint32_t)box_h, 1, 1};      // 64-byte rows either way
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(map, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 5, const_cast<void*>(base), dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? HUPR_OK : HUPR_ERR_CUDA;
}

// Weights [kd][kh][kw][cout][ld] seen as (cin, cout, kw, kd*3 + kh): a box of `box_kh` = 3 consecutive entries of the last dimension is the
// three kh taps of one (kd, kw) — one TMA instruction instead of three.
static int encode_halo_wgt_map(CUtensorMap* map, const void* base, int cin, int cout, int kd, int kw, int bn, int ld, int box_kh, int esize = 2) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return HUPR_ERR_CUDA;
    const cuuint64_t e = (cuuint64_t)esize;
    cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)cout, (cuuint64_t)kw, (cuuint64_t)kd * 3};
    cuuint64_t strides[3] = {(cuuint64_t)ld * e, (cuuint64_t)cout * ld * e, (cuuint64_t)kw * cout * ld * e};
    cuuint32_t box[4] = {(cuuint32_t)(64 / esize), (cuuint32_t)bn, 1, (cuuint32_t)box_kh};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(base), dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? HUPR_OK : HUPR_ERR_CUDA;
}

// How many 2-CTA clusters of the pair kernels can be co-resident on the current device (cached per device; 0 = clusters of this size
// cannot be scheduled here, e.g. a partitioned GPU: conv_halo_try then keeps the single-CTA kernels instead of failing).
static int halo_pair_clusters() {
    static int cached[kMaxDevices] = {};      // 0 = not queried yet, -1 = unavailable
    static bool configured[kMaxDevices] = {};
    const int dev = device_index();
    if (dev < 0 || dev >= kMaxDevices) return 0;
    if (cached[dev] == 0) {
        auto kernel = conv_halo_kernel<128, 3, true>;
        int n = 0;
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cfg.blockDim = dim3(kHaloThreads);
        cfg.dynamicSmemBytes = 232448;
        cfg.gridDim = dim3(2 * (device_sm_count() > 0 ? device_sm_count() : 1), 1, 1);
        if (ensure_smem_optin(kernel, 232448, configured) != HUPR_OK || cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n <= 0) {
            cudaGetLastError();      // a failed query is not an error of the convolution
            n = -1;
        }
        cached[dev] = n;
    }
    return cached[dev] > 0 ? cached[dev] : 0;
}

// CTA-pair launch (conv_halo_kernel<.., true>): clusters of two CTAs, as many as can be co-resident (a persistent kernel must not queue).
template <int BN, int NPROD>
static int launch_halo_pairs(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& a_x, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
                             const CUtensorMap& b_x, const ConvParams& p, const HaloGeom& g, int m_tiles, cudaStream_t stream) {
    static bool configured[kMaxDevices] = {};
    static int max_clusters[kMaxDevices] = {};
    const int smem_max = 232448;
    auto kernel = conv_halo_kernel<BN, NPROD, true>;
    if (int crc = ensure_smem_optin(kernel, smem_max, configured)) return crc;
    const int smem = g.stages * g.stage_bytes + 1024 + 256 + kHaloStatBytes;
    if (smem > smem_max) return HUPR_ERR_BAD_ARG;
    HaloGeom gg = g;
    gg.m_tiles = m_tiles;
    gg.n_tiles = p.cout / BN;
    const int pairs = gg.m_tiles * gg.n_tiles / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kHaloThreads);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int dev = device_index();
    if (dev < 0 || dev >= kMaxDevices) return HUPR_ERR_CUDA;
    if (max_clusters[dev] == 0) {
        cfg.gridDim = dim3(2 * device_sm_count(), 1, 1);
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n <= 0) return HUPR_ERR_CUDA;
        max_clusters[dev] = n;
    }
    int clusters = pairs < max_clusters[dev] ? pairs : max_clusters[dev];
    if (const char* cap = getenv("HUPR_HALO_GRID")) {
        const int c = atoi(cap) / 2;
        if (c > 0 && c < clusters) clusters = c;
    }
    cfg.gridDim = dim3(2 * clusters, 1, 1);
    if (pdl_enabled()) {
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.numAttrs = 2;
    }
    cudaLaunchKernelEx(&cfg, kernel, a_hi, a_lo, a_x, b_hi, b_lo, b_x, p, gg);
    note_launches(1);
    return cudaGetLastError() == cudaSuccess ? HUPR_OK : HUPR_ERR_CUDA;
}

template <int BN, int NPROD>
static int launch_halo(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& a_x, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
                       const CUtensorMap& b_x, const ConvParams& p, const HaloGeom& g, int m_tiles, cudaStream_t stream) {
    static bool configured[kMaxDevices] = {};
    const int smem_max = 232448;
    if (int crc = ensure_smem_optin(conv_halo_kernel<BN, NPROD, false>, smem_max, configured)) return crc;
    const int smem = g.stages * g.stage_bytes + 1024 + 256 + kHaloStatBytes;
    if (smem > smem_max) return HUPR_ERR_BAD_ARG;
    const int num_sms = device_sm_count();
    if (num_sms <= 0) return HUPR_ERR_CUDA;
    HaloGeom gg = g;
    gg.m_tiles = m_tiles;
    gg.n_tiles = p.cout / BN;
    const int total = gg.m_tiles * gg.n_tiles;
    int ctas = total < num_sms ? total : num_sms;
    if (const char* cap = getenv("HUPR_HALO_GRID")) {      // measurement switch: fewer persistent CTAs (per-SM vs chip-wide operand bandwidth)
        const int c = atoi(cap);
        if (c > 0 && c < ctas) ctas = c;
    }
    dim3 grid(ctas, 1, 1);
    launch_k(conv_halo_kernel<BN, NPROD, false>, grid, dim3(kHaloThreads), (size_t)smem, stream, a_hi, a_lo, a_x, b_hi, b_lo, b_x, p, gg);
    note_launches(1);
    return cudaGetLastError() == cudaSuccess ? HUPR_OK : HUPR_ERR_CUDA;
}

// Returns HUPR_OK after launching, or a positive value (1) if the shape is not handled here (caller falls through to the
// generic kernel).  Arguments were validated by hupr_conv_gemm.
// `probe`: no launch; returns 2 when the two-unit (NPROD == 2) kernel would take the descriptor, 1 otherwise.
int conv_halo_try(const hupr_conv_desc* d, const ConvParams& base, cudaStream_t stream, bool probe) {
    const bool three = d->a_lo != nullptr && d->w_lo != nullptr && d->nprod != 1;
    if (d->w_batched || d->k_split > 1 || d->w_k_off != 0) return 1;
    if (!three && getenv("HUPR_HALO1_OFF")) return 1;      // A/B switch: single-product convolutions on the generic kernel
    if (d->kh != 3 || d->ph != 1 || d->kw != 2 * d->pw + 1) return 1;
    if (d->w != 16 && d->w != 32 && d->w != 64) return 1;
    const int bw = d->w, bh = HM / bw;
    if (d->h % bh) return 1;
    const int d_out = base.d_out;
    const int m_tiles = d->n * d_out * (d->h / bh);
    const int bn = (d->cout % 128 == 0) ? 128 : 64;
    if (m_tiles * (d->cout / bn) < 120) return 1;      // too few CTAs to fill the machine: the 128-row generic tiles spread wider
    // contracted channels actually present in memory (the generic kernel lets cin run past ca into zero fill)
    int cin_eff = d->cin;
    if (d->a_ch_off + cin_eff > d->ca) cin_eff = d->ca - d->a_ch_off;
    if (cin_eff <= 0 || cin_eff % HK) return 1;

    ConvParams p = base;
    p.bw = bw; p.bh = bh; p.tiles_w = 1; p.tiles_h = d->h / bh;
    p.cin_blocks = cin_eff / HK;
    // CTA pairs (cta_group::2 MMAs, each CTA holds half of the weight rows): cout = 128 tiles, an even number of position tiles so that the
    // two CTAs of a pair always share their weight column tile; HUPR_HALO_SINGLE=1 is the A/B switch
    const bool pairs = m_tiles % 2 == 0 && !getenv("HUPR_HALO_SINGLE") && (probe || halo_pair_clusters() > 0);
    const bool pairs64 = pairs && bn == 64 && three;      // [w_hi | w_lo] form: a 64-row X tile and a 32-row Y tile per tap and CTA
    HaloGeom g;
    g.b_merged = 0;
    g.halo_rows = (bh + 2) * bw;
    g.a_plane_bytes = g.halo_rows * 64;
    g.b_tile_bytes = (pairs && !pairs64 ? bn / 2 : bn) * 64;
    g.nprod = three ? 3 : 1;
    g.stage_bytes = pairs64 ? 2 * g.a_plane_bytes + 3 * g.b_tile_bytes + 3 * (g.b_tile_bytes / 2)
                            : (three ? 2 : 1) * (g.a_plane_bytes + 3 * g.b_tile_bytes);
    g.stages = (232448 - 1024 - 256 - kHaloStatBytes) / g.stage_bytes;
    if (g.stages > 4) g.stages = 4;
    if (g.stages < 2) return 1;
    g.tap_stride_bytes = bw * 64;
    g.acc_stride_bytes = (bh / 2) * bw * 64;
    if (g.a_plane_bytes % 1024 || g.b_tile_bytes % 1024) return 1;

    const int taps = d->kd * d->kh * d->kw;
    const int w_ld = d->w_ld ? d->w_ld : d->cin;
    // two-unit arithmetic: cout = 128-column tiles only (the cout = 64 tiles are bound by shared-memory operand reads either way and keep
    // the [w_hi | w_lo] form), dense samples, no fused statistics (train mode never asks for it)
    const bool quant = three && d->nprod == 2 && bn == 128 && d->a_n_stride == 0 && !d->stats && d->ca % 32 == 0 && d->a_ch_off % 32 == 0 &&
                       w_ld % 32 == 0 && d->w_ch_off % 32 == 0 && !getenv("HUPR_QUANT_OFF");
    if (probe) return quant ? 2 : 1;
    if (quant && d->a_q16 && d->w_q16) {
        // the e4m3 planes are byte tensors with 2 * channels bytes per position / filter row ([values | residuals] per 32-channel block):
        // the same map builders with esize 1, doubled channel extents and a 64-byte box
        CUtensorMap a16, a8, b16, b8;
        int rc;
        if ((rc = encode_halo_act_map(&a16, d->a_q16, d->ca, d->w, d->h, d->d, d->n, bw, bh + 2, 0)) != HUPR_OK) return rc;
        if ((rc = encode_halo_act_map(&a8, d->a_q8, 2 * d->ca, d->w, d->h, d->d, d->n, bw, bh + 2, 0, 1)) != HUPR_OK) return rc;
        const char* w16 = static_cast<const char*>(d->w_q16) + (size_t)d->w_ch_off * 2;
        const char* w8 = static_cast<const char*>(d->w_q8) + (size_t)d->w_ch_off * 2;
        const int brows = pairs ? bn / 2 : bn;
        if ((rc = encode_halo_wgt_map(&b16, w16, d->cin, d->cout, d->kd, d->kw, brows, w_ld, 3)) != HUPR_OK) return rc;
        if ((rc = encode_halo_wgt_map(&b8, w8, 2 * d->cin, d->cout, d->kd, d->kw, brows, 2 * w_ld, 3, 1)) != HUPR_OK) return rc;
        g.nprod = 2;
        p.acc_scale = 1.0f / 65536.0f;
        if (pairs) return launch_halo_pairs<128, 2>(a16, a8, a16, b16, b8, b16, p, g, m_tiles, stream);
        return launch_halo<128, 2>(a16, a8, a16, b16, b8, b16, p, g, m_tiles, stream);
    }
    const __nv_bfloat16* w_hi = static_cast<const __nv_bfloat16*>(d->w_hi) + d->w_ch_off;
    const __nv_bfloat16* w_lo = three ? static_cast<const __nv_bfloat16*>(d->w_lo) + d->w_ch_off : w_hi;
    CUtensorMap a_hi, a_lo, b_hi, b_lo;
    int rc;
    if ((rc = encode_halo_act_map(&a_hi, d->a_hi, d->ca, d->w, d->h, d->d, d->n, bw, bh + 2, d->a_n_stride)) != HUPR_OK) return rc;
    if ((rc = encode_halo_act_map(&a_lo, three ? d->a_lo : d->a_hi, d->ca, d->w, d->h, d->d, d->n, bw, bh + 2, d->a_n_stride)) != HUPR_OK) return rc;
    g.b_merged = 0;
    if (pairs64) {
        CUtensorMap b_y;
        if ((rc = encode_halo_wgt_map(&b_hi, w_hi, d->cin, d->cout, d->kd, d->kw, 64, w_ld, 3)) != HUPR_OK) return rc;
        if ((rc = encode_halo_wgt_map(&b_lo, w_lo, d->cin, d->cout, d->kd, d->kw, 64, w_ld, 3)) != HUPR_OK) return rc;
        if ((rc = encode_halo_wgt_map(&b_y, w_hi, d->cin, d->cout, d->kd, d->kw, 32, w_ld, 3)) != HUPR_OK) return rc;
        return launch_halo_pairs<64, 3>(a_hi, a_lo, a_hi, b_hi, b_lo, b_y, p, g, m_tiles, stream);
    }
    if (three) {
        // both planes through ONE map when lo lies a 16-byte-multiple above hi in the address space (always true for the halves of one
        // allocation, SplitTensor.from_float; otherwise one box per tile and plane)
        const long long plane = (const char*)w_lo - (const char*)w_hi;
        if (plane > 0 && plane % 16 == 0 && plane < (1LL << 40) && !getenv("HUPR_HALO_B_UNMERGED")) {
            EncodeTiledFn fn = get_encode_fn();
            if (!fn) return HUPR_ERR_CUDA;
            cuuint64_t dims[5] = {(cuuint64_t)d->cin, (cuuint64_t)d->cout, 2, (cuuint64_t)d->kw, (cuuint64_t)d->kd * 3};
            cuuint64_t strides[4] = {(cuuint64_t)w_ld * 2, (cuuint64_t)plane, (cuuint64_t)d->cout * w_ld * 2, (cuuint64_t)d->kw * d->cout * w_ld * 2};
            cuuint32_t box[5] = {(cuuint32_t)HK, (cuuint32_t)(pairs ? bn / 2 : bn), 2, 1, 3};
            cuuint32_t estr[5] = {1, 1, 1, 1, 1};
            if (fn(&b_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<__nv_bfloat16*>(w_hi), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
                g.b_merged = 1;
                b_lo = b_hi;
            }
        }
        if (!g.b_merged) {
            if ((rc = encode_halo_wgt_map(&b_hi, w_hi, d->cin, d->cout, d->kd, d->kw, pairs ? bn / 2 : bn, w_ld, 1)) != HUPR_OK) return rc;
            if ((rc = encode_halo_wgt_map(&b_lo, w_lo, d->cin, d->cout, d->kd, d->kw, pairs ? bn / 2 : bn, w_ld, 1)) != HUPR_OK) return rc;
        }
    } else {
        if ((rc = encode_halo_wgt_map(&b_hi, w_hi, d->cin, d->cout, d->kd, d->kw, pairs ? bn / 2 : bn, w_ld, 3)) != HUPR_OK) return rc;
        b_lo = b_hi;
    }
    if (pairs) {
        if (three) return launch_halo_pairs<128, 3>(a_hi, a_lo, a_hi, b_hi, b_lo, b_hi, p, g, m_tiles, stream);
        return bn == 128 ? launch_halo_pairs<128, 1>(a_hi, a_lo, a_hi, b_hi, b_lo, b_hi, p, g, m_tiles, stream)
                         : launch_halo_pairs<64, 1>(a_hi, a_lo, a_hi, b_hi, b_lo, b_hi, p, g, m_tiles, stream);
    }
    if (three)
        return bn == 128 ? launch_halo<128, 3>(a_hi, a_lo, a_hi, b_hi, b_lo, b_hi, p, g, m_tiles, stream)
                         : launch_halo<64, 3>(a_hi, a_lo, a_hi, b_hi, b_lo, b_hi, p, g, m_tiles, stream);
    return bn == 128 ? launch_halo<128, 1>(a_hi, a_lo, a_hi, b_hi, b_lo, b_hi, p, g, m_tiles, stream)
                     : launch_halo<64, 1>(a_hi, a_lo, a_hi, b_hi, b_lo, b_hi, p, g, m_tiles, stream);
}

}  // namespace hupr
