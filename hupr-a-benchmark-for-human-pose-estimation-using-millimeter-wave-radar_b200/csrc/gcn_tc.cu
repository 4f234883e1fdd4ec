// Pose-refinement GCN around the tensor-core GEMM (sm_100a).
//
// PRGCN (/root/reference/models/gcn_networks.py:47-64) is three layers  Y = W[1024x1024] . (X . A[14x14]) + bias[1024x14]  on
// node features X[b] in R^{1024 x 14}.  The fp32 SIMT version (heads.cu, hupr_prgcn_fwd) is FMA-bound at batch 32 (0.94 GFLOP per
// layer); here the W-contraction runs on hupr_conv_gemm in the TRANSPOSED layout
//     Yt[(b, j), q] = sum_p St[(b, j), p] * W[q, p]        rows (b, j) = "positions", p = contracted "channels", W K-major as stored
// with bias as a residual tensor, ReLU as the epilogue slope, and these three small kernels for the stages in between:
//   hupr_gcn_nodes  heatmap = sigmoid(logits); bilinear x0.5 (align_corners); St0 = (X . A)^T as bf16 hi/lo rows [(b, j)][1024]
//   hupr_gcn_mix    St_{l+1}[(b, j'), q] = sum_j A[j, j'] * Yt_l[(b, j), q]                       (the next layer's  X . A)
//   hupr_gcn_heads  Yt_3 rows are the [B][14][32][32] maps: bilinear x2 (align_corners) + sigmoid -> [B][14][64][64]
#include "common.cuh"
#include "split.cuh"

namespace hupr {

constexpr int kGJ = 14;
constexpr int kGNodes = 1024;

__device__ __forceinline__ float gsigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(256)
gcn_nodes_kernel(const float* __restrict__ logits, int ld, const float* __restrict__ adj, float* __restrict__ heatmap,
                 __nv_bfloat16* __restrict__ s_hi, __nv_bfloat16* __restrict__ s_lo, int batch) {
    pdl_wait();                  // programmatic dependent launch: the predecessor's writes are visible from here on (common.cuh)
    pdl_launch_dependents();     // the successor may start its prologue now; it waits the same way before touching memory
    __shared__ float sA[kGJ * kGJ];
    if (threadIdx.x < kGJ * kGJ) sA[threadIdx.x] = __ldg(adj + threadIdx.x);
    __syncthreads();
    const int gid = blockIdx.x * 256 + threadIdx.x;
    if (gid < batch * 4096) {
        const int b = gid >> 12, pos = gid & 4095;
        const float* src = logits + (size_t)gid * ld;
#pragma unroll
        for (int k = 0; k < kGJ; ++k) heatmap[((size_t)b * kGJ + k) * 4096 + pos] = gsigmoid(__ldg(src + k));
    }
    if (gid < batch * kGNodes) {
        const int b = gid >> 10, node = gid & 1023;
        const int oh = node >> 5, ow = node & 31;
        const float scale = 63.0f / 31.0f;
        const float fh = scale * oh, fw = scale * ow;
        const int h0 = (int)fh, w0 = (int)fw;
        const int h1 = h0 + (h0 < 63), w1 = w0 + (w0 < 63);
        const float lh1 = fh - h0, lw1 = fw - w0, lh0 = 1.f - lh1, lw0 = 1.f - lw1;
        const float* base = logits + (size_t)b * 4096 * ld;
        float x[kGJ];
#pragma unroll
        for (int k = 0; k < kGJ; ++k) {
            const float v00 = __ldg(base + (size_t)(h0 * 64 + w0) * ld + k), v01 = __ldg(base + (size_t)(h0 * 64 + w1) * ld + k);
            const float v10 = __ldg(base + (size_t)(h1 * 64 + w0) * ld + k), v11 = __ldg(base + (size_t)(h1 * 64 + w1) * ld + k);
            x[k] = lh0 * (lw0 * v00 + lw1 * v01) + lh1 * (lw0 * v10 + lw1 * v11);
        }
#pragma unroll
        for (int j = 0; j < kGJ; ++j) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < kGJ; ++k) s = fmaf(x[k], sA[k * kGJ + j], s);
            const __nv_bfloat16 h = __float2bfloat16_rn(s);
            const size_t o = (size_t)(b * kGJ + j) * kGNodes + node;
            s_hi[o] = h;
            s_lo[o] = __float2bfloat16_rn(s - __bfloat162float(h));
        }
    }
}

// One thread per (sample, pair of feature columns q, q+1).
__global__ void __launch_bounds__(256)
gcn_mix_kernel(const uint32_t* __restrict__ y_hi, const uint32_t* __restrict__ y_lo, const float* __restrict__ adj,
               uint32_t* __restrict__ s_hi, uint32_t* __restrict__ s_lo, int batch) {
    pdl_wait();                  // programmatic dependent launch: the predecessor's writes are visible from here on (common.cuh)
    pdl_launch_dependents();     // the successor may start its prologue now; it waits the same way before touching memory
    __shared__ float sA[kGJ * kGJ];
    if (threadIdx.x < kGJ * kGJ) sA[threadIdx.x] = __ldg(adj + threadIdx.x);
    __syncthreads();
    const int gid = blockIdx.x * 256 + threadIdx.x;
    if (gid >= batch * (kGNodes / 2)) return;
    const int b = gid / (kGNodes / 2), qp = gid % (kGNodes / 2);
    float y0[kGJ], y1[kGJ];
#pragma unroll
    for (int j = 0; j < kGJ; ++j) {
        const size_t o = (size_t)(b * kGJ + j) * (kGNodes / 2) + qp;
        const uint32_t h = __ldg(y_hi + o), l = __ldg(y_lo + o);
        y0[j] = bf16_lo_f(h) + bf16_lo_f(l);
        y1[j] = bf16_hi_f(h) + bf16_hi_f(l);
    }
#pragma unroll
    for (int jp = 0; jp < kGJ; ++jp) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int j = 0; j < kGJ; ++j) {
            a0 = fmaf(y0[j], sA[j * kGJ + jp], a0);
            a1 = fmaf(y1[j], sA[j * kGJ + jp], a1);
        }
        uint32_t h, l;
        split2(a0, a1, h, l);
        const size_t o = (size_t)(b * kGJ + jp) * (kGNodes / 2) + qp;
        s_hi[o] = h;
        s_lo[o] = l;
    }
}

__global__ void __launch_bounds__(256)
gcn_heads_kernel(const float* __restrict__ y, int y_ld, float* __restrict__ gcn_heatmap, int batch) {
    pdl_wait();                  // programmatic dependent launch: the predecessor's writes are visible from here on (common.cuh)
    pdl_launch_dependents();     // the successor may start its prologue now; it waits the same way before touching memory
    const int gid = blockIdx.x * 256 + threadIdx.x;
    if (gid >= batch * kGJ * 4096) return;
    const int pos = gid & 4095, bj = gid >> 12;
    const int oh = pos >> 6, ow = pos & 63;
    const float scale = 31.0f / 63.0f;
    const float fh = scale * oh, fw = scale * ow;
    const int h0 = (int)fh, w0 = (int)fw;
    const int h1 = h0 + (h0 < 31), w1 = w0 + (w0 < 31);
    const float lh1 = fh - h0, lw1 = fw - w0, lh0 = 1.f - lh1, lw0 = 1.f - lw1;
    const float* m = y + (size_t)bj * y_ld;
    const float v00 = __ldg(m + h0 * 32 + w0), v01 = __ldg(m + h0 * 32 + w1), v10 = __ldg(m + h1 * 32 + w0), v11 = __ldg(m + h1 * 32 + w1);
    gcn_heatmap[gid] = gsigmoid(lh0 * (lw0 * v00 + lw1 * v01) + lh1 * (lw0 * v10 + lw1 * v11));
}

static int gcn_check_sm100() {
    return device_check_sm100();      // cached per device (capi.cu)
}

}  // namespace hupr

using namespace hupr;

extern "C" int hupr_gcn_nodes(const float* logits, int ld, const float* adj, float* heatmap, void* s_hi, void* s_lo, int batch, void* stream) {
    if (batch < 0) return HUPR_ERR_BAD_ARG;
    if (batch == 0) return HUPR_OK;
    if (!logits || !adj || !heatmap || !s_hi || !s_lo || ld < kGJ) return HUPR_ERR_BAD_ARG;
    int rc = gcn_check_sm100();
    if (rc != HUPR_OK) return rc;
    launch_k(gcn_nodes_kernel, dim3((batch * 4096 + 255) / 256), dim3(256), (size_t)(0), (cudaStream_t)stream, logits, ld, adj, heatmap, (__nv_bfloat16*)s_hi,
                                                                                   (__nv_bfloat16*)s_lo, batch);
    note_launches(1);
    return cudaGetLastError() == cudaSuccess ? HUPR_OK : HUPR_ERR_CUDA;
}

extern "C" int hupr_gcn_mix(const void* y_hi, const void* y_lo, const float* adj, void* s_hi, void* s_lo, int batch, void* stream) {
    if (batch < 0) return HUPR_ERR_BAD_ARG;
    if (batch == 0) return HUPR_OK;
    if (!y_hi || !y_lo || !adj || !s_hi || !s_lo) return HUPR_ERR_BAD_ARG;
    if (((uintptr_t)y_hi | (uintptr_t)y_lo | (uintptr_t)s_hi | (uintptr_t)s_lo) & 3) return HUPR_ERR_ALIGNMENT;
    int rc = gcn_check_sm100();
    if (rc != HUPR_OK) return rc;
    launch_k(gcn_mix_kernel, dim3((batch * (kGNodes / 2) + 255) / 256), dim3(256), (size_t)(0), (cudaStream_t)stream, 
        (const uint32_t*)y_hi, (const uint32_t*)y_lo, adj, (uint32_t*)s_hi, (uint32_t*)s_lo, batch);
    note_launches(1);
    return cudaGetLastError() == cudaSuccess ? HUPR_OK : HUPR_ERR_CUDA;
}

extern "C" int hupr_gcn_heads(const float* y, int y_ld, float* gcn_heatmap, int batch, void* stream) {
    if (batch < 0) return HUPR_ERR_BAD_ARG;
    if (batch == 0) return HUPR_OK;
    if (!y || !gcn_heatmap || y_ld < kGNodes) return HUPR_ERR_BAD_ARG;
    int rc = gcn_check_sm100();
    if (rc != HUPR_OK) return rc;
    launch_k(gcn_heads_kernel, dim3((batch * kGJ * 4096 + 255) / 256), dim3(256), (size_t)(0), (cudaStream_t)stream, y, y_ld, gcn_heatmap, batch);
    note_launches(1);
    return cudaGetLastError() == cudaSuccess ? HUPR_OK : HUPR_ERR_CUDA;
}
