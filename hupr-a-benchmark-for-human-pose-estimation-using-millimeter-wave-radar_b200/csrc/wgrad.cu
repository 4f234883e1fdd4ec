// Weight gradient of a stride-1 convolution straight from the channels-last tensors (tcgen05 + TMEM + TMA, sm_100a).
//
// Backward of every dense contraction of MSCSA-PRGCN with respect to its filter (what autograd derives for
// /root/reference/models/layers.py:24-32,45-63,116-123,195-210 during loss.backward(), /root/reference/tools/run.py:78):
//   dW[tap][ci][co] = sum_{n, d, h, w}  X[n, d + kd - pd, h + kh - ph, w + kw - pw, ci] * dY[n, d, h, w, co]        (zero outside X)
//
// GEMM view: the contracted axis is the POSITION axis; both operands are channels-last, i.e. for the tensor core they are
// MN-major (the 64 channels of a position are the contiguous 128 bytes, consecutive positions are consecutive 128-byte rows).
// The UMMA shared-memory descriptors describe exactly that (SWIZZLE_128B, MN-major canonical layout: 8-row groups 1024 B apart
// along K, 64-channel atoms LBO bytes apart along M/N), so the TMA boxes the forward convolution already uses are consumed
// without any transposed / position-major copy, and a filter tap is again nothing but a TMA coordinate offset with
// out-of-bounds zero fill as the padding.
//
//   * K block = 64 consecutive output positions (one [bh x bw] patch of one depth slice);
//   * A operand (TMEM lanes, M = 128) = two 64-channel "atoms" of the X side, an atom being (filter tap, 64-channel block of cin):
//     with cin = 64 a lane block pairs two taps, with cin >= 128 two channel blocks of one tap;
//   * B operand (N = BN columns of cout) = the dY box, loaded once per K block and shared by all the CTA's accumulator slots;
//   * a CTA owns 512 / BN accumulator slots (all 512 TMEM columns), a column tile of cout and a contiguous range of K blocks
//     (split-K); partial sums are added to the fp32 result with vector reductions (red.global.add.v4.f32);
//   * fp32-equivalent arithmetic as everywhere else: hi/lo bf16 planes, three products per k-step;
//   * warp 0 = TMA producer, warp 1 = single-thread tcgen05.mma issuer, warps 2..5 = epilogue.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "hupr_internal.h"
#include "tc.cuh"

namespace hupr {

constexpr int WG_KB = 64;                 // positions per K block
constexpr int WG_ATOM = WG_KB * 128;      // bytes of one atom plane: 64 positions x 64 bf16 channels
constexpr int WG_THREADS = 192;

template <int BN, int NPROD>
struct WgCfg {
    static constexpr int kPlanes = NPROD == 3 ? 2 : 1;           // single-product mode stages the hi planes only: twice the stages in flight
    static constexpr int kYAtoms = BN / 64;
    static constexpr int kYBytes = kPlanes * kYAtoms * WG_ATOM;  // hi atoms | lo atoms
    static constexpr int kABytes = kPlanes * 2 * WG_ATOM;        // hi atom 0, hi atom 1 | lo atom 0, lo atom 1
    static constexpr int kSlots = 512 / BN;                      // accumulator slots of [128 lanes x BN columns]
    static constexpr int kStagesFit = (224 * 1024 - 2 * kYBytes) / kABytes;
    static constexpr int kStages = kStagesFit > 12 ? 12 : kStagesFit;
    static constexpr int kSmemBytes = 2 * kYBytes + kStages * kABytes + 1024 /*align*/ + 256 /*barriers*/ + 256 /*atom table*/;
    static_assert(kStages >= 2 && 2 * kStages + 6 <= 32, "wgrad smem plan (32 barrier slots)");
};

struct WgParams {
    int n, d_out, h, w;
    int kh, kw, pd, ph, pw;
    int cin_blocks, x_ch_off, y_ch_off;
    int bw, bh, tiles_w, tiles_h;
    int atoms, slots, groups, n_tiles;
    int kblocks, kb_per_split, k_split;   // kblocks: K blocks of the whole tensor, or of ONE sample in batched mode
    float* out;
    int out_ld;
    long long out_batch_stride;            // batched mode: the samples are independent problems, dw[sample] = out + sample * stride
    int nprod;                             // 3: hi/lo planes, three products per k-step; 1: hi planes only (lo regions of a stage stay unused)
};

// MN-major, SWIZZLE_128B operand: rows (K) of 128 B = 64 channels; 8-row groups SBO = 1024 B apart; 64-channel atoms LBO apart.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int BN, int NPROD>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo,
             const __grid_constant__ CUtensorMap tmY_hi, const __grid_constant__ CUtensorMap tmY_lo, const WgParams p) {
    using Cfg = WgCfg<BN, NPROD>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sm_y = smem;                                   // 2 buffers of kYBytes
    uint8_t* sm_a = smem + 2 * Cfg::kYBytes;                // kStages stages of kABytes
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm_a + Cfg::kStages * Cfg::kABytes);
    int4* atom_tab = reinterpret_cast<int4*>(reinterpret_cast<uint8_t*>(bars) + 256);      // [2 * kSlots] (channel, dw, dh, dd) of the CTA's atoms
    uint64_t* a_full = bars;
    uint64_t* a_empty = bars + Cfg::kStages;
    uint64_t* y_full = bars + 2 * Cfg::kStages;             // [2]
    uint64_t* y_empty = bars + 2 * Cfg::kStages + 2;        // [2]
    uint64_t* acc_full = bars + 2 * Cfg::kStages + 4;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::kStages + 5);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr bool three = NPROD == 3;

    // work item: (sample [batched mode], split-K slice, column tile of cout, group of accumulator slots)
    int item = blockIdx.x;
    const int group = item % p.groups; item /= p.groups;
    const int n0 = (item % p.n_tiles) * BN; item /= p.n_tiles;
    const int z = item % p.k_split;
    const int sample = item / p.k_split;               // 0 unless batched
    const int kb_begin = z * p.kb_per_split;
    const int kb_end = min(p.kblocks, kb_begin + p.kb_per_split);
    const int slot0 = group * Cfg::kSlots;
    const int nslots = min(Cfg::kSlots, p.slots - slot0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&y_full[s], 1); mbar_init(&y_empty[s], 1); }
        mbar_init(acc_full, 1);
        fence_mbar_init();
        prefetch_tmap(&tmX_hi); prefetch_tmap(&tmX_lo); prefetch_tmap(&tmY_hi); prefetch_tmap(&tmY_lo);
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
            // The CTA's atoms (filter tap, 64-channel block) do not depend on the K block: decode them once (the integer divisions of
            // the decode were ~250 cycles per stage against 128..384 tensor cycles of work per stage) ...
            for (int j = 0; j < 2 * nslots; ++j) {
                int atom = slot0 * 2 + j;
                if (atom >= p.atoms) atom = p.atoms - 1;      // odd atom count: lanes 64..127 of the last slot repeat a real atom (ignored)
                const int tap = atom / p.cin_blocks, cb = atom - tap * p.cin_blocks;
                const int tkw = tap % p.kw, tkh = (tap / p.kw) % p.kh, tkd = tap / (p.kw * p.kh);
                atom_tab[j] = make_int4(p.x_ch_off + cb * 64, tkw - p.pw, tkh - p.ph, tkd - p.pd);
            }
            // ... and walk the K blocks (w tile fastest, then h tile, depth, sample) with carry counters instead of divisions
            int tw, th, od, nn;
            {
                int t = kb_begin;
                tw = t % p.tiles_w; t /= p.tiles_w;
                th = t % p.tiles_h; t /= p.tiles_h;
                od = t % p.d_out;
                nn = t / p.d_out;
            }
            int it = 0;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                const int w0 = tw * p.bw, h0 = th * p.bh, n = sample + nn;
                const int i = kb - kb_begin, yb = i & 1;
                mbar_wait(&y_empty[yb], (uint32_t)(((i >> 1) & 1) ^ 1));
                mbar_expect_tx(&y_full[yb], Cfg::kYBytes);
                uint8_t* ys = sm_y + yb * Cfg::kYBytes;
#pragma unroll
                for (int j = 0; j < Cfg::kYAtoms; ++j) {
                    tma_load_5d(ys + j * WG_ATOM, &tmY_hi, &y_full[yb], p.y_ch_off + n0 + 64 * j, w0, h0, od, n);
                    if (three) tma_load_5d(ys + (Cfg::kYAtoms + j) * WG_ATOM, &tmY_lo, &y_full[yb], p.y_ch_off + n0 + 64 * j, w0, h0, od, n);
                }
                for (int sl = 0; sl < nslots; ++sl, ++it) {
                    const int s = it % Cfg::kStages;
                    mbar_wait(&a_empty[s], (uint32_t)(((it / Cfg::kStages) & 1) ^ 1));
                    mbar_expect_tx(&a_full[s], Cfg::kABytes);
                    uint8_t* as = sm_a + s * Cfg::kABytes;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int4 at = atom_tab[2 * sl + half];
                        tma_load_5d(as + half * WG_ATOM, &tmX_hi, &a_full[s], at.x, w0 + at.y, h0 + at.z, od + at.w, n);
                        if (three) tma_load_5d(as + (2 + half) * WG_ATOM, &tmX_lo, &a_full[s], at.x, w0 + at.y, h0 + at.z, od + at.w, n);
                    }
                }
                if (++tw == p.tiles_w) {
                    tw = 0;
                    if (++th == p.tiles_h) {
                        th = 0;
                        if (++od == p.d_out) { od = 0; ++nn; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (one thread) =================
        if (elect_one()) {
            // D = f32, A = B = bf16, both MN-major (bits 15, 16), N = BN, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(128 >> 4) << 24);
            int it = 0;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                const int i = kb - kb_begin, yb = i & 1;
                mbar_wait(&y_full[yb], (uint32_t)((i >> 1) & 1));
                const uint32_t ys = smem_u32(sm_y + yb * Cfg::kYBytes);
                for (int sl = 0; sl < nslots; ++sl, ++it) {
                    const int s = it % Cfg::kStages;
                    mbar_wait(&a_full[s], (uint32_t)((it / Cfg::kStages) & 1));
                    tc_fence_after();
                    const uint32_t as = smem_u32(sm_a + s * Cfg::kABytes);
                    const uint32_t tacc = tmem_base + (uint32_t)(sl * BN);
#pragma unroll
                    for (int k = 0; k < WG_KB / 16; ++k) {
                        const uint32_t koff = (uint32_t)(k * 16 * 128);        // 16 positions = two 8-row groups
                        const uint64_t da_hi = make_smem_desc_mn(as + koff, WG_ATOM);
                        const uint64_t da_lo = make_smem_desc_mn(as + 2 * WG_ATOM + koff, WG_ATOM);
                        const uint64_t dy_hi = make_smem_desc_mn(ys + koff, WG_ATOM);
                        const uint64_t dy_lo = make_smem_desc_mn(ys + Cfg::kYAtoms * WG_ATOM + koff, WG_ATOM);
                        const uint32_t acc_flag = (kb != kb_begin || k != 0) ? 1u : 0u;
                        if (three) {
                            umma_bf16(tacc, da_lo, dy_hi, idesc, acc_flag);
                            umma_bf16(tacc, da_hi, dy_lo, idesc, 1u);
                            umma_bf16(tacc, da_hi, dy_hi, idesc, 1u);
                        } else {
                            umma_bf16(tacc, da_hi, dy_hi, idesc, acc_flag);
                        }
                    }
                    tc_commit(&a_empty[s]);
                }
                tc_commit(&y_empty[yb]);
            }
            tc_commit(acc_full);
        }
    } else {
        // ================= epilogue warps (2..5): TMEM lanes 32*(warp%4) .. +31 =================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        mbar_wait(acc_full, 0);
        tc_fence_after();
        for (int sl = 0; sl < nslots; ++sl) {
            const int atom = (slot0 + sl) * 2 + (row >> 6);
            const bool valid = atom < p.atoms;
            float* dst = p.out + (size_t)sample * p.out_batch_stride + ((size_t)atom * 64 + (row & 63)) * p.out_ld + n0;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t acc[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sl * BN + c * 32), acc);
                if (valid) {
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        red_add_v4(dst + c * 32 + g * 4, __uint_as_float(acc[4 * g]), __uint_as_float(acc[4 * g + 1]),
                                   __uint_as_float(acc[4 * g + 2]), __uint_as_float(acc[4 * g + 3]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ---- host side -----------------------------------------------------------------------------------
static int encode_pos_map(CUtensorMap* map, const void* base, int c, int w, int h, int d, int n, int bw, int bh) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return HUPR_ERR_CUDA;
    cuuint64_t dims[5] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)d, (cuuint64_t)n};
    cuuint64_t strides[4] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2, (cuuint64_t)d * h * w * c * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? HUPR_OK : HUPR_ERR_CUDA;
}

template <int BN, int NPROD>
static int launch_wgrad(const CUtensorMap& x_hi, const CUtensorMap& x_lo, const CUtensorMap& y_hi, const CUtensorMap& y_lo, WgParams p,
                        int cout, int problems, int num_sms, cudaStream_t stream) {
    using Cfg = WgCfg<BN, NPROD>;
    static bool configured[kMaxDevices] = {};
    if (int crc = ensure_smem_optin(wgrad_kernel<BN, NPROD>, Cfg::kSmemBytes, configured)) return crc;
    p.groups = (p.slots + Cfg::kSlots - 1) / Cfg::kSlots;
    p.n_tiles = cout / BN;
    const long long tiles = (long long)p.groups * p.n_tiles * problems;
    int ks = (int)(num_sms / tiles);                       // one wave of CTAs when the tile count allows it
    if (ks < 1) ks = 1;
    if (ks > p.kblocks) ks = p.kblocks;
    p.kb_per_split = (p.kblocks + ks - 1) / ks;
    p.k_split = (p.kblocks + p.kb_per_split - 1) / p.kb_per_split;
    const long long items = tiles * p.k_split;
    if (items > 2147483647LL) return HUPR_ERR_BAD_ARG;
    wgrad_kernel<BN, NPROD><<<(unsigned)items, WG_THREADS, Cfg::kSmemBytes, stream>>>(x_hi, x_lo, y_hi, y_lo, p);
    note_launches(1);
    return cudaGetLastError() == cudaSuccess ? HUPR_OK : HUPR_ERR_CUDA;
}

}  // namespace hupr

extern "C" int hupr_conv_wgrad(const hupr_wgrad_desc* d, void* stream) {
    using namespace hupr;
    if (!d || !d->x_hi || !d->dy_hi || !d->dw) return HUPR_ERR_BAD_ARG;
    if (d->nprod != 0 && d->nprod != 1 && d->nprod != 3) return HUPR_ERR_BAD_ARG;
    const bool three = d->nprod != 1;
    if (three && (!d->x_lo || !d->dy_lo)) return HUPR_ERR_BAD_ARG;
    if (d->n <= 0 || d->d <= 0 || d->h <= 0 || d->w <= 0) return HUPR_ERR_BAD_ARG;
    if (d->kd <= 0 || d->kh <= 0 || d->kw <= 0 || d->pd < 0) return HUPR_ERR_BAD_ARG;
    if (d->kh != 2 * d->ph + 1 || d->kw != 2 * d->pw + 1) return HUPR_ERR_BAD_ARG;            // H, W are 'same' convolutions
    const int d_out = d->d + 2 * d->pd - d->kd + 1;
    if (d_out <= 0) return HUPR_ERR_BAD_ARG;
    if (d->cin <= 0 || d->cin % 64 || d->cout <= 0 || d->cout % 64) return HUPR_ERR_BAD_ARG;
    if (d->cx % 8 || d->x_ch_off % 8 || d->x_ch_off < 0 || d->x_ch_off + d->cin > ((d->cx + 63) / 64) * 64) return HUPR_ERR_BAD_ARG;
    if (d->cy % 8 || d->y_ch_off % 8 || d->y_ch_off < 0 || d->y_ch_off + d->cout > ((d->cy + 63) / 64) * 64) return HUPR_ERR_BAD_ARG;
    if (d->dw_ld % 4 || d->dw_ld < d->cout) return HUPR_ERR_BAD_ARG;
    const int bw = d->w < WG_KB ? d->w : WG_KB;
    if (WG_KB % bw || d->w % bw) return HUPR_ERR_BAD_ARG;
    const int bh = WG_KB / bw;
    if (d->h % bh) return HUPR_ERR_BAD_ARG;
    const uintptr_t align_or = (uintptr_t)d->x_hi | (uintptr_t)(three ? d->x_lo : nullptr) | (uintptr_t)d->dy_hi | (uintptr_t)(three ? d->dy_lo : nullptr) | (uintptr_t)d->dw;
    if (align_or & 15) return HUPR_ERR_ALIGNMENT;
    if (int arch_rc = device_check_sm100()) return arch_rc;
    const int num_sms = device_sm_count();
    const int problems = d->batched ? d->n : 1;
    if (d->batched && (d->dw_batch_stride < 0 || d->dw_batch_stride % 4)) return HUPR_ERR_BAD_ARG;
    const long long kblocks = (long long)(d->batched ? 1 : d->n) * d_out * (d->h / bh) * (d->w / bw);
    if (kblocks > 2147483647LL) return HUPR_ERR_BAD_ARG;

    WgParams p;
    p.n = d->n; p.d_out = d_out; p.h = d->h; p.w = d->w;
    p.kh = d->kh; p.kw = d->kw; p.pd = d->pd; p.ph = d->ph; p.pw = d->pw;
    p.cin_blocks = d->cin / 64; p.x_ch_off = d->x_ch_off; p.y_ch_off = d->y_ch_off;
    p.bw = bw; p.bh = bh; p.tiles_w = d->w / bw; p.tiles_h = d->h / bh;
    p.atoms = d->kd * d->kh * d->kw * p.cin_blocks;
    p.slots = (p.atoms + 1) / 2;
    p.groups = 0; p.n_tiles = 0;
    p.kblocks = (int)kblocks; p.kb_per_split = 0; p.k_split = 0;
    p.out = d->dw; p.out_ld = d->dw_ld; p.out_batch_stride = d->batched ? d->dw_batch_stride : 0;
    p.nprod = three ? 3 : 1;

    CUtensorMap x_hi, x_lo, y_hi, y_lo;
    int rc;
    if ((rc = encode_pos_map(&x_hi, d->x_hi, d->cx, d->w, d->h, d->d, d->n, bw, bh)) != HUPR_OK) return rc;
    if ((rc = encode_pos_map(&x_lo, three ? d->x_lo : d->x_hi, d->cx, d->w, d->h, d->d, d->n, bw, bh)) != HUPR_OK) return rc;
    if ((rc = encode_pos_map(&y_hi, d->dy_hi, d->cy, d->w, d->h, d_out, d->n, bw, bh)) != HUPR_OK) return rc;
    if ((rc = encode_pos_map(&y_lo, three ? d->dy_lo : d->dy_hi, d->cy, d->w, d->h, d_out, d->n, bw, bh)) != HUPR_OK) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (three) {
        if (d->cout % 256 == 0) return launch_wgrad<256, 3>(x_hi, x_lo, y_hi, y_lo, p, d->cout, problems, num_sms, s);
        if (d->cout % 128 == 0) return launch_wgrad<128, 3>(x_hi, x_lo, y_hi, y_lo, p, d->cout, problems, num_sms, s);
        return launch_wgrad<64, 3>(x_hi, x_lo, y_hi, y_lo, p, d->cout, problems, num_sms, s);
    }
    if (d->cout % 256 == 0) return launch_wgrad<256, 1>(x_hi, x_lo, y_hi, y_lo, p, d->cout, problems, num_sms, s);
    if (d->cout % 128 == 0) return launch_wgrad<128, 1>(x_hi, x_lo, y_hi, y_lo, p, d->cout, problems, num_sms, s);
    return launch_wgrad<64, 1>(x_hi, x_lo, y_hi, y_lo, p, d->cout, problems, num_sms, s);
}
