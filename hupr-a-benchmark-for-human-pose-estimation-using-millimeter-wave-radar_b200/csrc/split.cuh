// bf16 hi/lo split helpers shared by the HBM-bound kernels (value = hi + lo, both bf16; see conv_gemm.cu).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace hupr {

__device__ __forceinline__ uint32_t bf16_bits(float v) { return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v)); }
__device__ __forceinline__ float bf16_lo_f(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16_hi_f(uint32_t packed) { return __uint_as_float(packed & 0xFFFF0000u); }

// Split two floats into packed (hi, lo) bf16x2 words: element 0 in the low half (one cvt.rn.bf16x2.f32 per plane).
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xFFFF0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(ra, rb);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// 8 channels (16 B of hi + 16 B of lo) -> 8 floats.  lo may be null (single-bf16 tensors).
__device__ __forceinline__ void load8(const __nv_bfloat16* hi, const __nv_bfloat16* lo, float (&v)[8]) {
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(hi));
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = bf16_lo_f(hw[i]);
        v[2 * i + 1] = bf16_hi_f(hw[i]);
    }
    if (lo) {
        const uint4 l = __ldg(reinterpret_cast<const uint4*>(lo));
        const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] += bf16_lo_f(lw[i]);
            v[2 * i + 1] += bf16_hi_f(lw[i]);
        }
    }
}

__device__ __forceinline__ void store8(__nv_bfloat16* hi, __nv_bfloat16* lo, const float (&v)[8]) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split2(v[2 * i], v[2 * i + 1], h[i], l[i]);
    *reinterpret_cast<uint4*>(hi) = make_uint4(h[0], h[1], h[2], h[3]);
    if (lo) *reinterpret_cast<uint4*>(lo) = make_uint4(l[0], l[1], l[2], l[3]);
}

}  // namespace hupr
