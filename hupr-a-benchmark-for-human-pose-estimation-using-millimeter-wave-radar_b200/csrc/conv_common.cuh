// Shared between the implicit-GEMM convolution kernels: launch parameters and the fused epilogue
// (scale/shift -> + residual -> PReLU-style slope -> fp32 and/or bf16 hi/lo stores) for 32 accumulator columns of one row.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

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
};

// acc: 32 consecutive accumulator columns (channels ch0 .. ch0+31) of output position `pos`.
__device__ __forceinline__ void conv_epilogue32(const ConvParams& p, const uint32_t (&acc)[32], size_t pos, int ch0) {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        float x = __uint_as_float(acc[j]);
        const float sc = p.scale ? __ldg(p.scale + ch0 + j) : 1.0f;
        const float sh = p.shift ? __ldg(p.shift + ch0 + j) : 0.0f;
        v[j] = fmaf(x, sc, sh);
    }
    if (p.r_hi) {
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
        for (int j = 0; j < 16; ++j) {
            const __nv_bfloat16 h0b = __float2bfloat16_rn(v[2 * j]), h1b = __float2bfloat16_rn(v[2 * j + 1]);
            const __nv_bfloat16 l0b = __float2bfloat16_rn(v[2 * j] - __bfloat162float(h0b));
            const __nv_bfloat16 l1b = __float2bfloat16_rn(v[2 * j + 1] - __bfloat162float(h1b));
            hi[j] = (uint32_t)__bfloat16_as_ushort(h0b) | ((uint32_t)__bfloat16_as_ushort(h1b) << 16);
            lo[j] = (uint32_t)__bfloat16_as_ushort(l0b) | ((uint32_t)__bfloat16_as_ushort(l1b) << 16);
        }
        uint4* dh = reinterpret_cast<uint4*>(p.o_hi + pos * p.o_ld + p.o_ch_off + ch0);
#pragma unroll
        for (int g = 0; g < 4; ++g) dh[g] = make_uint4(hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3]);
        if (p.o_lo) {
            uint4* dl = reinterpret_cast<uint4*>(p.o_lo + pos * p.o_ld + p.o_ch_off + ch0);
#pragma unroll
            for (int g = 0; g < 4; ++g) dl[g] = make_uint4(lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3]);
        }
    }
}

}  // namespace hupr
