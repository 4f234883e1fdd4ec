// bf16 hi/lo split helpers shared by the HBM-bound kernels (value = hi + lo, both bf16; see conv_gemm.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
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

// Operand planes of the two-unit convolution arithmetic (hupr_conv_desc.nprod == 2, include/hupr_b200.h): 8 values -> 8 fp16 of
// x * s16 (saturating), 8 e4m3 of x * s8 and 8 e4m3 of (x - x16) * s8l, where x16 is the fp16 plane's value.  The scales are powers of two.
struct QuantScales {
    float s16, inv16, s8, s8l;
};
__device__ __forceinline__ QuantScales quant_scales(bool is_weight) {
    QuantScales q;
    q.s16 = is_weight ? 16384.f : 4.f;          // fp16 plane: activations up to 16 376, weights up to 3.99 before saturation
    q.inv16 = is_weight ? 1.f / 16384.f : 1.f / 4.f;
    q.s8 = is_weight ? 16.f : 2.f;
    q.s8l = is_weight ? 32768.f : 4096.f;
    return q;
}
__device__ __forceinline__ uint32_t e4m3x2(float a, float b) {      // a in the low byte
    return (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
}
// Returns true when a value left the fp16 plane's range (|x * s16| > 65504: the plane saturates and the result is no longer the value).
__device__ __forceinline__ bool quant8(const float (&v)[8], const QuantScales& q, uint4& h16, uint2& q8, uint2& q8l) {
    uint32_t h[4], a8[4], l8[4];
    float top = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a = v[2 * i], b = v[2 * i + 1];
        top = fmaxf(top, fmaxf(fabsf(a), fabsf(b)));
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(b * q.s16), "f"(a * q.s16));      // a in the low half; saturates to +-65504
        const __half2 hh = *reinterpret_cast<const __half2*>(&h[i]);
        const float ra = a - __low2float(hh) * q.inv16, rb = b - __high2float(hh) * q.inv16;
        a8[i] = e4m3x2(a * q.s8, b * q.s8);
        l8[i] = e4m3x2(ra * q.s8l, rb * q.s8l);
    }
    h16 = make_uint4(h[0], h[1], h[2], h[3]);
    q8 = make_uint2(a8[0] | (a8[1] << 16), a8[2] | (a8[3] << 16));
    q8l = make_uint2(l8[0] | (l8[1] << 16), l8[2] | (l8[3] << 16));
    return !(top * q.s16 <= 65504.f);      // also true for NaN
}
// Saturation report of the quantised planes: a device counter that is only ever incremented (callers read and reset it).
__device__ __forceinline__ void quant_report(int* flag, bool saturated) {
    if (saturated && flag) atomicAdd(flag, 1);
}

}  // namespace hupr
