// Shared between the implicit-GEMM convolution kernels: launch parameters and the fused epilogue
// (scale/shift -> + residual -> PReLU-style slope -> fp32 and/or bf16 hi/lo stores) for 32 accumulator columns of one row.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "split.cuh"

namespace hupr {

struct ConvParams {
    int n, d_in, h, w, d_out;
    int kd, kh, kw, pd, ph, pw;
    int cin_blocks, a_ch_off, w_batched;
    int bw, bh, tiles_w, tiles_h;
    int cout;
    const float* scale;
    const float* shift;
    const float* slope;
    const __nv_bfloat16* r_hi;
    const __nv_bfloat16* r_lo;
    int r_ld, r_ch_off;
    __nv_bfloat16* o_hi;
    __nv_bfloat16* o_lo;
    int o_ld, o_ch_off;
    float* o_f32;
    int o_f32_ld;
    int w_k_off;                   // added to the B operand's contracted-axis coordinate (shifted correlation; out of range = zero fill)
    int kb_per_split, k_split, atomic;
    int m_tiles, n_tiles;          // 128-position tiles x BN-column tiles (x k_split slices), walked persistently
    const float* row_vec;          // per-position vector (row_mode 1: exp(acc - v), 2: r * (acc - v))
    int row_mode;
    float* coop_ws;                // cooperative split-K: fp32 partial tiles [k_split][tiles][128][BN] ...
    int* coop_counters;            // ... and one arrival counter per (tile, 32-row quarter, column half); zero before and after every launch
    int coop;
    size_t coop_ws_bytes;
    int n_fastest;                 // work-item order: column tile fastest (wide outputs) instead of row tile fastest
    double* stats;                 // optional per-output-channel sums of the epilogue values: stats[ch] += sum v, stats[stats_ld + ch] += sum v^2
    int stats_ld;                  // (train-mode BatchNorm batch statistics fused into the producing convolution, layers.py:45-53)
    float acc_scale;               // exact power of two applied to the accumulator first (2^-16 for the two-unit arithmetic, else 1)
    __half* o_q16;                 // optional second form of the output: activation operand planes of the two-unit arithmetic
    uint8_t* o_q8;                 // ([positions][o_ld] at o_ch_off like o_hi; e4m3 plane [positions][o_ld / 32][2][32]; hupr_conv_desc.o_q*)
    int* q_sat;                    // optional device counter: incremented when a value written to o_q16 left the fp16 plane's range
    int st256;                     // 1: every output plane row segment is 32-byte aligned -> 256-bit stores (one full sector per lane)
};

// 32-byte store (STG.256, sm_100): a thread that owns one output row writes whole 32-byte sectors instead of two half-sector pieces in two
// instructions — the thread-per-row epilogue's scattered 16-byte stores were what the L1/L2 write path choked on.
__device__ __forceinline__ void st_global_256(void* ptr, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5, uint32_t a6,
                                              uint32_t a7) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5),
                 "r"(a6), "r"(a7)
                 : "memory");
}

// Column sums over a warp: every lane holds 32 values (one row of a 32-row x 32-column block); after five exchange rounds
// (recursive halving, 31 shuffles) lane j holds the sum of column j over the 32 rows.
__device__ __forceinline__ float warp_column_sums32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = upper ? v[i] : v[i + off];
            const float keep = upper ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

// Fused BatchNorm statistics: `wstat` = one warp's running column sums in shared memory, [sum v | sum v^2] x `width` columns; lane j owns
// columns j, j + 32, ...  stat_add folds a 32-row x 32-column block in (31 + 31 shuffles), stat_flush adds the totals to global memory.
__device__ __forceinline__ void stat_add(float* wstat32, int width, const float (&v)[32], int lane) {
    float a[32], b[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) { a[j] = v[j]; b[j] = v[j] * v[j]; }
    const float s1 = warp_column_sums32(a, lane), s2 = warp_column_sums32(b, lane);
    wstat32[lane] += s1;
    wstat32[width + lane] += s2;
}
// Direct form (generic kernel: few tiles per CTA, column tile changes from item to item): the block's column sums go straight to global memory.
__device__ __forceinline__ void stat_add_global(const ConvParams& p, const float (&v)[32], int ch0, int lane) {
    float a[32], b[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) { a[j] = v[j]; b[j] = v[j] * v[j]; }
    const float s1 = warp_column_sums32(a, lane), s2 = warp_column_sums32(b, lane);
    atomicAdd(p.stats + ch0 + lane, (double)s1);
    atomicAdd(p.stats + p.stats_ld + ch0 + lane, (double)s2);
}
__device__ __forceinline__ void stat_flush(const ConvParams& p, float* wstat, int width, int ch0, int lane) {
    for (int j = lane; j < width; j += 32) {
        atomicAdd(p.stats + ch0 + j, (double)wstat[j]);
        atomicAdd(p.stats + p.stats_ld + ch0 + j, (double)wstat[width + j]);
        wstat[j] = 0.f;
        wstat[width + j] = 0.f;
    }
}

// acc: 32 consecutive accumulator columns (channels ch0 .. ch0+31) of output position `pos` -> epilogue values v (everything but the stores).
// rv = row_vec[pos] (loaded by the caller ahead of time; ignored unless row_mode).
__device__ __forceinline__ void conv_epilogue_values(const ConvParams& p, const uint32_t (&acc)[32], size_t pos, int ch0, float (&v)[32], float rv) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        float x = __uint_as_float(acc[j]);
        const float sc = (p.scale ? __ldg(p.scale + ch0 + j) : 1.0f) * p.acc_scale;     // power-of-two factor: exact
        const float sh = p.shift ? __ldg(p.shift + ch0 + j) : 0.0f;
        v[j] = fmaf(x, sc, sh);
    }
    if (p.row_mode) {
        if (p.row_mode == 1) {
            const float kLog2e = 1.4426950408889634f;
            const float nr = -rv * kLog2e;
#pragma unroll
            for (int j = 0; j < 32; ++j) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(v[j]) : "f"(fmaf(v[j], kLog2e, nr)));
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] -= rv;
        }
    }
    if (p.r_hi && p.row_mode == 2) {          // multiplicative operand: out = r * (acc - row_vec)
        const __nv_bfloat16* rh = p.r_hi + pos * p.r_ld + p.r_ch_off + ch0;
        const __nv_bfloat16* rl = p.r_lo ? p.r_lo + pos * p.r_ld + p.r_ch_off + ch0 : nullptr;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            float r8[8];
            load8(rh + g * 8, rl ? rl + g * 8 : nullptr, r8);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[g * 8 + e] *= r8[e];
        }
    } else if (p.r_hi) {
        const uint4* rh = reinterpret_cast<const uint4*>(p.r_hi + pos * p.r_ld + p.r_ch_off + ch0);
        const uint4* rl = p.r_lo ? reinterpret_cast<const uint4*>(p.r_lo + pos * p.r_ld + p.r_ch_off + ch0) : nullptr;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const uint4 a = __ldg(rh + g);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                v[g * 8 + 2 * e] += __uint_as_float(aw[e] << 16);
                v[g * 8 + 2 * e + 1] += __uint_as_float(aw[e] & 0xFFFF0000u);
            }
            if (rl) {
                const uint4 b = __ldg(rl + g);
                const uint32_t bw_[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    v[g * 8 + 2 * e] += __uint_as_float(bw_[e] << 16);
                    v[g * 8 + 2 * e + 1] += __uint_as_float(bw_[e] & 0xFFFF0000u);
                }
            }
        }
    }
    if (p.slope) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float sl = __ldg(p.slope + ch0 + j);
            v[j] = v[j] > 0.0f ? v[j] : v[j] * sl;
        }
    }
}

// Stores of 32 epilogue values of one row (fp32 and/or hi/lo bf16; split-K partial sums are added atomically).
__device__ __forceinline__ void conv_epilogue_store(const ConvParams& p, const float (&v)[32], size_t pos, int ch0) {
    if (p.atomic) {     // split-K partial sum (no scale / shift / residual / activation on this path)
        float* dst = p.o_f32 + pos * p.o_f32_ld + ch0;
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(dst + j, v[j]);
        return;
    }
    if (p.o_f32) {
        float4* dst = reinterpret_cast<float4*>(p.o_f32 + pos * p.o_f32_ld + ch0);
#pragma unroll
        for (int g = 0; g < 8; ++g) dst[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
    }
    if (p.o_hi) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);      // same roundings, one cvt.rn.bf16x2 per plane
        uint4* dh = reinterpret_cast<uint4*>(p.o_hi + pos * p.o_ld + p.o_ch_off + ch0);
        uint4* dl = p.o_lo ? reinterpret_cast<uint4*>(p.o_lo + pos * p.o_ld + p.o_ch_off + ch0) : nullptr;
        if (p.st256) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                st_global_256(dh + 2 * g, hi[8 * g], hi[8 * g + 1], hi[8 * g + 2], hi[8 * g + 3], hi[8 * g + 4], hi[8 * g + 5], hi[8 * g + 6], hi[8 * g + 7]);
                if (dl) st_global_256(dl + 2 * g, lo[8 * g], lo[8 * g + 1], lo[8 * g + 2], lo[8 * g + 3], lo[8 * g + 4], lo[8 * g + 5], lo[8 * g + 6], lo[8 * g + 7]);
            }
        } else {
#pragma unroll
            for (int g = 0; g < 4; ++g) dh[g] = make_uint4(hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3]);
            if (dl) {
#pragma unroll
                for (int g = 0; g < 4; ++g) dl[g] = make_uint4(lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3]);
            }
        }
    }
    if (p.o_q16 && p.st256) {
        const QuantScales q = quant_scales(false);
        const size_t at = pos * p.o_ld + p.o_ch_off + ch0;
        uint4 h[4];
        uint2 a8[4], l8[4];
        bool sat = false;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const float v8[8] = {v[8 * g], v[8 * g + 1], v[8 * g + 2], v[8 * g + 3], v[8 * g + 4], v[8 * g + 5], v[8 * g + 6], v[8 * g + 7]};
            sat |= quant8(v8, q, h[g], a8[g], l8[g]);
        }
        quant_report(p.q_sat, sat);
        st_global_256(p.o_q16 + at, h[0].x, h[0].y, h[0].z, h[0].w, h[1].x, h[1].y, h[1].z, h[1].w);
        st_global_256(p.o_q16 + at + 16, h[2].x, h[2].y, h[2].z, h[2].w, h[3].x, h[3].y, h[3].z, h[3].w);
        uint8_t* d8 = p.o_q8 + 2 * at;          // ch0 and o_ch_off are multiples of 32: this chunk's 64-byte block [values | residuals]
        st_global_256(d8, a8[0].x, a8[0].y, a8[1].x, a8[1].y, a8[2].x, a8[2].y, a8[3].x, a8[3].y);
        st_global_256(d8 + 32, l8[0].x, l8[0].y, l8[1].x, l8[1].y, l8[2].x, l8[2].y, l8[3].x, l8[3].y);
    } else if (p.o_q16) {
        const QuantScales q = quant_scales(false);
        const size_t at = pos * p.o_ld + p.o_ch_off + ch0;
        uint4* d16 = reinterpret_cast<uint4*>(p.o_q16 + at);
        uint2* d8 = reinterpret_cast<uint2*>(p.o_q8 + 2 * at);
        uint2* d8l = d8 + 4;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const float v8[8] = {v[8 * g], v[8 * g + 1], v[8 * g + 2], v[8 * g + 3], v[8 * g + 4], v[8 * g + 5], v[8 * g + 6], v[8 * g + 7]};
            uint4 h;
            uint2 a8, l8;
            quant_report(p.q_sat, quant8(v8, q, h, a8, l8));
            d16[g] = h;
            d8[g] = a8;
            d8l[g] = l8;
        }
    }
}

__device__ __forceinline__ void conv_epilogue32(const ConvParams& p, const uint32_t (&acc)[32], size_t pos, int ch0, float rv) {
    float v[32];
    conv_epilogue_values(p, acc, pos, ch0, v, rv);
    conv_epilogue_store(p, v, pos, ch0);
}

// Same, also folding the row's values into the running column sums of its chunk (fused BatchNorm statistics; warp-uniform branch).
__device__ __forceinline__ void conv_epilogue32(const ConvParams& p, const uint32_t (&acc)[32], size_t pos, int ch0, float rv, float* wstat32,
                                                int width, int lane) {
    float v[32];
    conv_epilogue_values(p, acc, pos, ch0, v, rv);
    conv_epilogue_store(p, v, pos, ch0);
    if (p.stats) stat_add(wstat32, width, v, lane);
}

// Same with the direct global-memory form of the statistics.
__device__ __forceinline__ void conv_epilogue32_stats(const ConvParams& p, const uint32_t (&acc)[32], size_t pos, int ch0, float rv, int lane) {
    float v[32];
    conv_epilogue_values(p, acc, pos, ch0, v, rv);
    conv_epilogue_store(p, v, pos, ch0);
    if (p.stats) stat_add_global(p, v, ch0, lane);
}

__device__ __forceinline__ void conv_epilogue32(const ConvParams& p, const uint32_t (&acc)[32], size_t pos, int ch0) {
    conv_epilogue32(p, acc, pos, ch0, p.row_mode ? __ldg(p.row_vec + pos) : 0.f);
}

}  // namespace hupr
