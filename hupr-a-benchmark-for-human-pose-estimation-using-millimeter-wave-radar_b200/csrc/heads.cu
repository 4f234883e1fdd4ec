// Heatmap head, pose-refinement GCN and heatmap loss (sm_100a, fp32 SIMT — these stages are HBM / weight-stream bound).
//   hupr_prgcn_fwd          /root/reference/models/gcn_networks.py:6-64 (PRGCN, GCN_layers), adjacency models/layers.py:97-112,
//                           plus heatmap = sigmoid(logits) of HuPRNet.forward (models/networks.py:40)
//   hupr_keypoints_argmax   /root/reference/misc/metrics.py:10-38 (get_max_preds)
//   hupr_heatmap_loss_fwd   /root/reference/misc/losses.py:23-48 + misc/utils.py:6-65 (generateTarget): on-device Gaussian targets,
//                           2 x BCE(mean) and the target argmax coordinates
#include "common.cuh"

namespace hupr {

constexpr int kJ = 14;          // keypoints
constexpr int kNodes = 1024;    // 32 x 32 GCN feature rows
constexpr int kHm = 64;         // heatmap side

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---- stage 0: heatmap = sigmoid(logits) (NCHW out) ; bilinear x0.5 (align_corners) ; support0 = x . A ----------------
// logits: fp32 channels-last [B][4096][ld]; sup: [1024][ncol] with column b*14 + j.
__global__ void __launch_bounds__(256)
gcn_pre_kernel(const float* __restrict__ logits, int ld, const float* __restrict__ adj, float* __restrict__ heatmap,
               float* __restrict__ sup, int batch, int ncol) {
    __shared__ float sA[kJ * kJ];
    if (threadIdx.x < kJ * kJ) sA[threadIdx.x] = __ldg(adj + threadIdx.x);
    __syncthreads();
    const int gid = blockIdx.x * 256 + threadIdx.x;
    // part 1: sigmoid of every logit, written NCHW [B][14][64][64]
    if (gid < batch * 4096) {
        const int b = gid >> 12, pos = gid & 4095;
        const float* src = logits + (size_t)gid * ld;
#pragma unroll
        for (int k = 0; k < kJ; ++k) heatmap[((size_t)b * kJ + k) * 4096 + pos] = sigmoidf_(__ldg(src + k));
    }
    // part 2: one thread per (b, node)
    if (gid < batch * kNodes) {
        const int b = gid >> 10, node = gid & 1023;
        const int oh = node >> 5, ow = node & 31;
        const float scale = 63.0f / 31.0f;
        const float fh = scale * oh, fw = scale * ow;
        const int h0 = (int)fh, w0 = (int)fw;
        const int h1 = h0 + (h0 < 63), w1 = w0 + (w0 < 63);
        const float lh1 = fh - h0, lw1 = fw - w0, lh0 = 1.f - lh1, lw0 = 1.f - lw1;
        const float* base = logits + (size_t)b * 4096 * ld;
        float x[kJ];
#pragma unroll
        for (int k = 0; k < kJ; ++k) {
            const float v00 = __ldg(base + (size_t)(h0 * 64 + w0) * ld + k), v01 = __ldg(base + (size_t)(h0 * 64 + w1) * ld + k);
            const float v10 = __ldg(base + (size_t)(h1 * 64 + w0) * ld + k), v11 = __ldg(base + (size_t)(h1 * 64 + w1) * ld + k);
            x[k] = lh0 * (lw0 * v00 + lw1 * v01) + lh1 * (lw0 * v10 + lw1 * v11);
        }
        float* dst = sup + (size_t)node * ncol + b * kJ;
#pragma unroll
        for (int j = 0; j < kJ; ++j) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < kJ; ++k) s = fmaf(x[k], sA[k * kJ + j], s);
            dst[j] = s;
        }
    }
}

// ---- GCN layer: Y = W[1024x1024] . S[1024 x ncol] + bias[q][j]; optional ReLU; optional right-multiply by A for the next layer ---
// Tile: 16 rows x 56 columns (4 samples x 14 joints), K chunks of 32.  112 active threads of 128: thread (tq, tn) owns rows 2tq..2tq+1,
// columns 4tn..4tn+3.
constexpr int kGcnTQ = 16, kGcnTN = 56, kGcnTK = 32;

__global__ void __launch_bounds__(128)
gcn_layer_kernel(const float* __restrict__ W, const float* __restrict__ S, const float* __restrict__ bias, const float* __restrict__ adj,
                 float* __restrict__ out, int ncol, int relu, int mul_adj) {
    __shared__ float sW[kGcnTQ][kGcnTK + 1];
    __shared__ float sS[kGcnTK][kGcnTN];
    __shared__ float sY[kGcnTQ][kGcnTN];
    __shared__ float sA[kJ * kJ];
    const int tid = threadIdx.x;
    const int q0 = blockIdx.x * kGcnTQ, n0 = blockIdx.y * kGcnTN;
    const int ncols_here = min(kGcnTN, ncol - n0);
    for (int i = tid; i < kJ * kJ; i += 128) sA[i] = __ldg(adj + i);
    const bool active = tid < 112;
    const int tq = tid / 14, tn = tid % 14;
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    for (int k0 = 0; k0 < kNodes; k0 += kGcnTK) {
        __syncthreads();
        for (int i = tid; i < kGcnTQ * kGcnTK; i += 128) {
            const int r = i / kGcnTK, c = i % kGcnTK;
            sW[r][c] = __ldg(W + (size_t)(q0 + r) * kNodes + k0 + c);
        }
        for (int i = tid; i < kGcnTK * kGcnTN; i += 128) {
            const int r = i / kGcnTN, c = i % kGcnTN;
            sS[r][c] = (c < ncols_here) ? __ldg(S + (size_t)(k0 + r) * ncol + n0 + c) : 0.f;
        }
        __syncthreads();
        if (active) {
#pragma unroll 8
            for (int k = 0; k < kGcnTK; ++k) {
                const float w0 = sW[2 * tq][k], w1 = sW[2 * tq + 1][k];
                const float4 s = *reinterpret_cast<const float4*>(&sS[k][4 * tn]);
                acc[0][0] = fmaf(w0, s.x, acc[0][0]); acc[0][1] = fmaf(w0, s.y, acc[0][1]);
                acc[0][2] = fmaf(w0, s.z, acc[0][2]); acc[0][3] = fmaf(w0, s.w, acc[0][3]);
                acc[1][0] = fmaf(w1, s.x, acc[1][0]); acc[1][1] = fmaf(w1, s.y, acc[1][1]);
                acc[1][2] = fmaf(w1, s.z, acc[1][2]); acc[1][3] = fmaf(w1, s.w, acc[1][3]);
            }
        }
    }
    if (active) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int col = 4 * tn + c;
                float y = acc[r][c] + __ldg(bias + (size_t)(q0 + 2 * tq + r) * kJ + col % kJ);
                if (relu) y = fmaxf(y, 0.f);
                sY[2 * tq + r][col] = y;
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < kGcnTQ * kGcnTN; i += 128) {
        const int r = i / kGcnTN, col = i % kGcnTN;
        if (col >= ncols_here) continue;
        float y;
        if (mul_adj) {
            const int g = col / kJ, j = col % kJ;
            y = 0.f;
#pragma unroll
            for (int k = 0; k < kJ; ++k) y = fmaf(sY[r][g * kJ + k], sA[k * kJ + j], y);
        } else {
            y = sY[r][col];
        }
        out[(size_t)(q0 + r) * ncol + n0 + col] = y;
    }
}

// ---- stage 4: [1024][ncol] -> [B][14][32][32] -> bilinear x2 (align_corners) -> sigmoid -> [B][14][64][64] -------------------
__global__ void __launch_bounds__(256)
gcn_post_kernel(const float* __restrict__ y, int ncol, float* __restrict__ gcn_heatmap, int batch) {
    const int gid = blockIdx.x * 256 + threadIdx.x;
    if (gid >= batch * kJ * 4096) return;
    const int pos = gid & 4095, bj = gid >> 12;       // bj = b*14 + j = column
    const int oh = pos >> 6, ow = pos & 63;
    const float scale = 31.0f / 63.0f;
    const float fh = scale * oh, fw = scale * ow;
    const int h0 = (int)fh, w0 = (int)fw;
    const int h1 = h0 + (h0 < 31), w1 = w0 + (w0 < 31);
    const float lh1 = fh - h0, lw1 = fw - w0, lh0 = 1.f - lh1, lw0 = 1.f - lw1;
    const float v00 = __ldg(y + (size_t)(h0 * 32 + w0) * ncol + bj), v01 = __ldg(y + (size_t)(h0 * 32 + w1) * ncol + bj);
    const float v10 = __ldg(y + (size_t)(h1 * 32 + w0) * ncol + bj), v11 = __ldg(y + (size_t)(h1 * 32 + w1) * ncol + bj);
    gcn_heatmap[gid] = sigmoidf_(lh0 * (lw0 * v00 + lw1 * v01) + lh1 * (lw0 * v10 + lw1 * v11));
}

// ---- argmax over each 64x64 map (first maximum wins, like numpy.argmax) --------------------------------------------------------
__global__ void __launch_bounds__(256)
argmax_kernel(const float* __restrict__ maps, float* __restrict__ preds, float* __restrict__ maxvals) {
    pdl_wait();                  // programmatic dependent launch: the predecessor's writes are visible from here on (common.cuh)
    pdl_launch_dependents();     // the successor may start its prologue now; it waits the same way before touching memory
    __shared__ float sv[8];
    __shared__ int si[8];
    const float* m = maps + (size_t)blockIdx.x * 4096;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < 4096; i += 256) {
        const float v = __ldg(m + i);
        if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            if (sv[w] > best || (sv[w] == best && si[w] < bi)) { best = sv[w]; bi = si[w]; }
        const float mask = best > 0.0f ? 1.0f : 0.0f;
        preds[blockIdx.x * 2] = (float)(bi & 63) * mask;
        preds[blockIdx.x * 2 + 1] = (float)(bi >> 6) * mask;
        if (maxvals) maxvals[blockIdx.x] = best;
    }
}

// ---- loss: targets synthesised on the fly -----------------------------------------------------------------------------------------
// joints: int64 [B][14][2] image-pixel coordinates.  mu = int(j / 4 + 0.5) (truncation toward zero, utils.py:37-38).
__device__ __forceinline__ bool joint_center(const long long* joints, int bk, int& mx, int& my) {
    mx = (int)((double)joints[bk * 2] / 4.0 + 0.5);
    my = (int)((double)joints[bk * 2 + 1] / 4.0 + 0.5);
    // utils.py:42: skipped when the 13x13 patch is entirely outside the map
    return !(mx - 6 >= kHm || my - 6 >= kHm || mx + 7 < 0 || my + 7 < 0);
}

__device__ __forceinline__ float target_value(bool valid, int mx, int my, int x, int y) {
    const int dx = x - mx, dy = y - my;
    if (!valid || dx < -6 || dx > 6 || dy < -6 || dy > 6) return 0.f;
    return expf(-(float)(dx * dx + dy * dy) / 8.0f);
}

__device__ __forceinline__ float bce_term(float p, float t) {
    const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(logf(1.0f - p), -100.f);   // torch clamps the logs at -100
    return -(t * lp + (1.0f - t) * lq);
}

// One CTA per (b, k) map: partial sums of both BCE terms; optional target / gt-argmax output.
__global__ void __launch_bounds__(256)
loss_partial_kernel(const float* __restrict__ heatmap, const float* __restrict__ gcn, const long long* __restrict__ joints,
                    double* __restrict__ partial /* [maps][2] */, float* __restrict__ targets, float* __restrict__ gt2d) {
    __shared__ double red[2][8];
    const int bk = blockIdx.x;
    int mx, my;
    const bool valid = joint_center(joints, bk, mx, my);
    const float* h = heatmap + (size_t)bk * 4096;
    const float* g = gcn + (size_t)bk * 4096;
    double s1 = 0.0, s2 = 0.0;
    for (int i = threadIdx.x; i < 4096; i += 256) {
        const float t = target_value(valid, mx, my, i & 63, i >> 6);
        if (targets) targets[(size_t)bk * 4096 + i] = t;
        s1 += (double)bce_term(__ldg(h + i), t);
        s2 += (double)bce_term(__ldg(g + i), t);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
        partial[bk * 2] = a;
        partial[bk * 2 + 1] = b;
        if (gt2d) {
            // argmax of the target map: the Gaussian centre if it lies inside the map, else the first maximum of the clipped patch
            float best = -INFINITY;
            int bi = 0;
            if (valid) {
                for (int y = max(my - 6, 0); y <= min(my + 6, kHm - 1); ++y)
                    for (int x = max(mx - 6, 0); x <= min(mx + 6, kHm - 1); ++x) {
                        const float t = target_value(true, mx, my, x, y);
                        if (t > best) { best = t; bi = y * 64 + x; }
                    }
            }
            const float mask = best > 0.0f ? 1.0f : 0.0f;
            gt2d[bk * 2] = (float)(bi & 63) * mask;
            gt2d[bk * 2 + 1] = (float)(bi >> 6) * mask;
        }
    }
}

__global__ void loss_final_kernel(const double* __restrict__ partial, int maps, float* __restrict__ out /* [loss, loss2, loss1] */) {
    __shared__ double red[2][32];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < maps; i += blockDim.x) { a += partial[i * 2]; b += partial[i * 2 + 1]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = 0.0; b = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += red[0][w]; b += red[1][w]; }
        const double n = (double)maps * 4096.0;
        const float l1 = (float)(a / n), l2 = (float)(b / n);
        out[0] = l1 + l2;    // lossDecay == -1 -> loss = loss1 + loss2 (losses.py:39-42)
        out[1] = l2;
        out[2] = l1;
    }
}

static int heads_check_sm100() {
    return device_check_sm100();      // cached per device (capi.cu)
}

}  // namespace hupr

using namespace hupr;

extern "C" size_t hupr_prgcn_workspace_bytes(int batch) {
    if (batch <= 0) return 0;
    return (size_t)2 * kNodes * ((size_t)batch * kJ) * sizeof(float);
}

extern "C" int hupr_prgcn_fwd(const float* logits, int ld, const float* const* weights, const float* const* biases, const float* adj,
                              void* workspace, size_t ws_bytes, float* heatmap, float* gcn_heatmap, int batch, void* stream) {
    if (batch < 0) return HUPR_ERR_BAD_ARG;
    if (batch == 0) return HUPR_OK;
    if (!logits || !weights || !biases || !adj || !workspace || !heatmap || !gcn_heatmap || ld < kJ) return HUPR_ERR_BAD_ARG;
    for (int l = 0; l < 3; ++l)
        if (!weights[l] || !biases[l]) return HUPR_ERR_BAD_ARG;
    if (ws_bytes < hupr_prgcn_workspace_bytes(batch)) return HUPR_ERR_WORKSPACE;
    if ((uintptr_t)workspace & 15) return HUPR_ERR_ALIGNMENT;
    int rc = heads_check_sm100();
    if (rc != HUPR_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int ncol = batch * kJ;
    float* buf0 = static_cast<float*>(workspace);
    float* buf1 = buf0 + (size_t)kNodes * ncol;
    gcn_pre_kernel<<<(batch * 4096 + 255) / 256, 256, 0, s>>>(logits, ld, adj, heatmap, buf0, batch, ncol);
    const dim3 grid(kNodes / kGcnTQ, (ncol + kGcnTN - 1) / kGcnTN);
    gcn_layer_kernel<<<grid, 128, 0, s>>>(weights[0], buf0, biases[0], adj, buf1, ncol, 1, 1);
    gcn_layer_kernel<<<grid, 128, 0, s>>>(weights[1], buf1, biases[1], adj, buf0, ncol, 1, 1);
    gcn_layer_kernel<<<grid, 128, 0, s>>>(weights[2], buf0, biases[2], adj, buf1, ncol, 0, 0);
    gcn_post_kernel<<<(batch * kJ * 4096 + 255) / 256, 256, 0, s>>>(buf1, ncol, gcn_heatmap, batch);
    note_launches(5);
    return cudaGetLastError() == cudaSuccess ? HUPR_OK : HUPR_ERR_CUDA;
}

extern "C" int hupr_keypoints_argmax(const float* maps, int n_maps, float* preds, float* maxvals, void* stream) {
    if (n_maps < 0) return HUPR_ERR_BAD_ARG;
    if (n_maps == 0) return HUPR_OK;
    if (!maps || !preds) return HUPR_ERR_BAD_ARG;
    int rc = heads_check_sm100();
    if (rc != HUPR_OK) return rc;
    launch_k(argmax_kernel, dim3(n_maps), dim3(256), (size_t)(0), (cudaStream_t)stream, maps, preds, maxvals);
    note_launches(1);
    return cudaGetLastError() == cudaSuccess ? HUPR_OK : HUPR_ERR_CUDA;
}

// Object keypoint similarity of one detection against one ground-truth pose per image (the reference's vendored COCOeval.computeOks,
// /root/reference/misc/cocoeval.py:192-236, for the HuPR case: every joint labelled visible).  float64 like the numpy original.
__global__ void oks_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const double* __restrict__ area,
                           const double* __restrict__ sigmas, int n, int k, double* __restrict__ oks, double* __restrict__ per_joint) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double denom = area[i] + 2.220446049250313e-16;      // np.spacing(1)
    double acc = 0.0;
    for (int j = 0; j < k; ++j) {
        const double dx = (double)pred[(i * k + j) * 2] - (double)gt[(i * k + j) * 2];
        const double dy = (double)pred[(i * k + j) * 2 + 1] - (double)gt[(i * k + j) * 2 + 1];
        const double var = (sigmas[j] * 2.0) * (sigmas[j] * 2.0);
        const double e = (dx * dx + dy * dy) / var / denom / 2.0;
        const double sim = exp(-e);
        if (per_joint) per_joint[i * k + j] = sim;
        acc += sim;
    }
    oks[i] = acc / (double)k;
}

extern "C" int hupr_keypoint_oks(const float* pred, const float* gt, const double* area, const double* sigmas, int n, int k, double* oks,
                                 double* per_joint, void* stream) {
    if (n < 0 || k <= 0) return HUPR_ERR_BAD_ARG;
    if (n == 0) return HUPR_OK;
    if (!pred || !gt || !area || !sigmas || !oks) return HUPR_ERR_BAD_ARG;
    int rc = heads_check_sm100();
    if (rc != HUPR_OK) return rc;
    oks_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(pred, gt, area, sigmas, n, k, oks, per_joint);
    note_launches(1);
    return cudaGetLastError() == cudaSuccess ? HUPR_OK : HUPR_ERR_CUDA;
}

extern "C" int hupr_heatmap_loss_fwd(const float* heatmap, const float* gcn_heatmap, const long long* joints, int batch,
                                     void* workspace, size_t ws_bytes, float* losses, float* targets, float* gt2d, void* stream) {
    if (batch < 0) return HUPR_ERR_BAD_ARG;
    if (batch == 0) return HUPR_OK;
    if (!heatmap || !gcn_heatmap || !joints || !workspace || !losses) return HUPR_ERR_BAD_ARG;
    const int maps = batch * kJ;
    if (ws_bytes < (size_t)maps * 2 * sizeof(double)) return HUPR_ERR_WORKSPACE;
    if ((uintptr_t)workspace & 7) return HUPR_ERR_ALIGNMENT;
    int rc = heads_check_sm100();
    if (rc != HUPR_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    loss_partial_kernel<<<maps, 256, 0, s>>>(heatmap, gcn_heatmap, joints, static_cast<double*>(workspace), targets, gt2d);
    loss_final_kernel<<<1, 256, 0, s>>>(static_cast<const double*>(workspace), maps, losses);
    note_launches(2);
    return cudaGetLastError() == cudaSuccess ? HUPR_OK : HUPR_ERR_CUDA;
}

// ================================================================================================================================
// Training building blocks (row a-17 of SURVEY.md §8; assembled into the whole backward pass by hupr_b200/training.py — DESIGN.md §3b).
//   hupr_heatmap_loss_bwd  d(loss1 + loss2)/d(logits) through BCE(mean) and the sigmoids  (reference: autograd of misc/losses.py:23-48)
//   hupr_adam_step         torch.optim.Adam with coupled L2 weight decay on a flat fp32 buffer (reference: tools/base.py:47)
// ================================================================================================================================
namespace hupr {

__device__ __forceinline__ float bce_sigmoid_grad(float p, float t, float inv_n) {
    // BCE backward: (p - t) / max(p (1 - p), 1e-12) / N ; sigmoid backward: * p (1 - p)
    const float pq = p * (1.0f - p);
    return (p - t) / fmaxf(pq, 1e-12f) * pq * inv_n;
}

// One CTA per (b, k) map.  d_heat_logits: channels-last [B][4096][ld] (the head conv's output layout); d_gcn_pre: [B][14][64][64].
__global__ void __launch_bounds__(256)
loss_bwd_kernel(const float* __restrict__ heatmap, const float* __restrict__ gcn, const long long* __restrict__ joints, int ld,
                float inv_n1, float inv_n2, float* __restrict__ d_heat_logits, float* __restrict__ d_gcn_pre) {
    const int bk = blockIdx.x;
    const int b = bk / kJ, k = bk % kJ;
    int mx, my;
    const bool valid = joint_center(joints, bk, mx, my);
    for (int i = threadIdx.x; i < 4096; i += 256) {
        const float t = target_value(valid, mx, my, i & 63, i >> 6);
        d_heat_logits[((size_t)b * 4096 + i) * ld + k] = bce_sigmoid_grad(__ldg(heatmap + (size_t)bk * 4096 + i), t, inv_n1);
        d_gcn_pre[(size_t)bk * 4096 + i] = bce_sigmoid_grad(__ldg(gcn + (size_t)bk * 4096 + i), t, inv_n2);
    }
}

// Sigmoid backward for caller-supplied output gradients (the autograd bridge of HuPRNet.forward in train() mode): g_* are dL/d(post-sigmoid
// maps) [B][14][64][64]; writes dL/d(pre-sigmoid) in the layouts loss_bwd_kernel uses.
__global__ void __launch_bounds__(256)
heatmap_bwd_kernel(const float* __restrict__ heatmap, const float* __restrict__ gcn, const float* __restrict__ g_heat,
                   const float* __restrict__ g_gcn, int ld, float* __restrict__ d_heat_logits, float* __restrict__ d_gcn_pre) {
    const int bk = blockIdx.x;
    const int b = bk / kJ, k = bk % kJ;
    for (int i = threadIdx.x; i < 4096; i += 256) {
        const size_t at = (size_t)bk * 4096 + i;
        const float h = __ldg(heatmap + at), s = __ldg(gcn + at);
        d_heat_logits[((size_t)b * 4096 + i) * ld + k] = g_heat ? __ldg(g_heat + at) * h * (1.0f - h) : 0.0f;
        d_gcn_pre[at] = g_gcn ? __ldg(g_gcn + at) * s * (1.0f - s) : 0.0f;
    }
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
            float step_size, float bc2_sqrt, float beta1, float beta2, float eps, float weight_decay, float lr, double bc1,
            const int* __restrict__ step_dev, const float* __restrict__ lr_dev) {
    if (step_dev) {      // step count lives on the device (CUDA-graph replays): derive the bias corrections here
        const double t = (double)__ldg(step_dev);
        bc1 = 1.0 - pow((double)beta1, t);
        bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, t));
    }
    if (lr_dev) lr = __ldg(lr_dev);      // learning rate lives on the device: a schedule keeps working under CUDA-graph replay
    if (step_dev || lr_dev) step_size = (float)((double)lr / bc1);
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float pi = p[i];
        const float gi = fmaf(weight_decay, pi, g[i]);            // coupled L2: grad += wd * param
        const float mi = m[i] + (1.0f - beta1) * (gi - m[i]);     // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = fmaf((1.0f - beta2) * gi, gi, v[i] * beta2);
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - step_size * (mi / denom);
    }
}

}  // namespace hupr

extern "C" int hupr_heatmap_bwd(const float* heatmap, const float* gcn_heatmap, const float* g_heat, const float* g_gcn, int batch, int ld,
                                float* d_heat_logits, float* d_gcn_pre, void* stream) {
    if (batch < 0) return HUPR_ERR_BAD_ARG;
    if (batch == 0) return HUPR_OK;
    if (!heatmap || !gcn_heatmap || !d_heat_logits || !d_gcn_pre || ld < kJ) return HUPR_ERR_BAD_ARG;
    int rc = heads_check_sm100();
    if (rc != HUPR_OK) return rc;
    heatmap_bwd_kernel<<<batch * kJ, 256, 0, (cudaStream_t)stream>>>(heatmap, gcn_heatmap, g_heat, g_gcn, ld, d_heat_logits, d_gcn_pre);
    note_launches(1);
    return cudaGetLastError() == cudaSuccess ? HUPR_OK : HUPR_ERR_CUDA;
}

extern "C" int hupr_heatmap_loss_bwd(const float* heatmap, const float* gcn_heatmap, const long long* joints, int batch, int ld,
                                     float w1, float w2, float* d_heat_logits, float* d_gcn_pre, void* stream) {
    if (batch < 0) return HUPR_ERR_BAD_ARG;
    if (batch == 0) return HUPR_OK;
    if (!heatmap || !gcn_heatmap || !joints || !d_heat_logits || !d_gcn_pre || ld < kJ) return HUPR_ERR_BAD_ARG;
    int rc = heads_check_sm100();
    if (rc != HUPR_OK) return rc;
    const float inv_n = 1.0f / ((float)batch * kJ * 4096.0f);
    loss_bwd_kernel<<<batch * kJ, 256, 0, (cudaStream_t)stream>>>(heatmap, gcn_heatmap, joints, ld, w1 * inv_n, w2 * inv_n, d_heat_logits, d_gcn_pre);
    note_launches(1);
    return cudaGetLastError() == cudaSuccess ? HUPR_OK : HUPR_ERR_CUDA;
}

extern "C" int hupr_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                              float beta2, float eps, float weight_decay, int step, const int* step_dev, const float* lr_dev, void* stream) {
    if (n < 0 || (step < 1 && !step_dev)) return HUPR_ERR_BAD_ARG;
    if (step < 1) step = 1;
    if (n == 0) return HUPR_OK;
    if (!params || !grads || !exp_avg || !exp_avg_sq) return HUPR_ERR_BAD_ARG;
    int rc = heads_check_sm100();
    if (rc != HUPR_OK) return rc;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    long long blocks = (n + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, (float)((double)lr / bc1),
                                                                    (float)sqrt(bc2), beta1, beta2, eps, weight_decay, lr, bc1, step_dev, lr_dev);
    note_launches(1);
    return cudaGetLastError() == cudaSuccess ? HUPR_OK : HUPR_ERR_CUDA;
}
