// HBM-bound kernels either side of the tensor-core contractions (sm_100a):
//   hupr_window_normalize  loader bridge: cascade cubes -> standardised VRDAE windows
//                          (/root/reference/datasets/base.py:13-24, datasets/dataset.py:120-150)
//   hupr_mnet_fwd          elevation mean + the (chirp, re/im)->(channel, depth) view + MNet
//                          (/root/reference/models/networks.py:23-33, models/chirp_networks.py:11-21)
//   hupr_resample_linear   bi/tri-linear align_corners=True resampling of split channels-last tensors
//                          (/root/reference/models/layers.py:84,89,199,204)
//   hupr_softmax_rows      softmax over the key axis of the attention logits (layers.py:131)
//   hupr_transpose_split   [n][s][c] -> [n][c][s] (V operand of the second attention matmul, layers.py:132)
#include "common.cuh"
#include "split.cuh"

namespace hupr {

constexpr int kPlaneCells = 64 * 64;          // (range, azimuth) cells per Doppler plane
constexpr int kPlaneF4 = kPlaneCells * 8 / 2; // float4 (= 2 complex) per Doppler plane
constexpr int kNormThreads = 512;

// Reduce 4 per-thread values over all threads that share (tid & 3); result broadcast to every thread.
__device__ __forceinline__ void reduce_by_quad(float (&v)[4], float* scratch /* [16 warps][4][4] */) {
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();   // scratch may still be read from a previous reduction
    if (lane < 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) scratch[(warp * 4 + lane) * 4 + k] = v[k];
    }
    __syncthreads();
    const int q = lane & 3;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kNormThreads / 32; ++w) s += scratch[(w * 4 + q) * 4 + k];
        v[k] = s;
    }
}

// One CTA per (window slot, kept Doppler row).  The 256 KiB plane is read three times (mean, centred variance, write);
// passes 2 and 3 hit L2.  Thread t owns float4 index t + 512*k, i.e. a fixed elevation pair 2*(t&3), 2*(t&3)+1.
__global__ void __launch_bounds__(kNormThreads)
window_normalize_kernel(const float4* __restrict__ cube, const int* __restrict__ slot_fs, float* __restrict__ vrdae) {
    pdl_wait();                  // programmatic dependent launch: the predecessor's writes are visible from here on (common.cuh)
    pdl_launch_dependents();     // the successor may start its prologue now; it waits the same way before touching memory
    __shared__ float scratch[(kNormThreads / 32) * 16];
    const int c = blockIdx.x;                       // kept Doppler row 0..7 -> cube row 4 + c (dataset.py:145)
    const int slot = blockIdx.y;
    const int fs = __ldg(slot_fs + slot);
    const float4* plane = cube + ((size_t)fs * 16 + 4 + c) * kPlaneF4;
    const int tid = threadIdx.x;

    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int k = 0; k < kPlaneF4 / kNormThreads; ++k) {
        const float4 v = __ldg(plane + tid + k * kNormThreads);
        s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
    }
    reduce_by_quad(s, scratch);
    float mean[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) mean[k] = s[k] * (1.0f / kPlaneCells);

    float q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int k = 0; k < kPlaneF4 / kNormThreads; ++k) {
        const float4 v = __ldg(plane + tid + k * kNormThreads);
        const float d0 = v.x - mean[0], d1 = v.y - mean[1], d2 = v.z - mean[2], d3 = v.w - mean[3];
        q[0] = fmaf(d0, d0, q[0]); q[1] = fmaf(d1, d1, q[1]); q[2] = fmaf(d2, d2, q[2]); q[3] = fmaf(d3, d3, q[3]);
    }
    reduce_by_quad(q, scratch);
    float rstd[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) rstd[k] = 1.0f / sqrtf(q[k] * (1.0f / (kPlaneCells - 1)));   // unbiased std (base.py:23)

    float2* out_re = reinterpret_cast<float2*>(vrdae + ((size_t)(slot * 8 + c) * 2 + 0) * (kPlaneCells * 8));
    float2* out_im = reinterpret_cast<float2*>(vrdae + ((size_t)(slot * 8 + c) * 2 + 1) * (kPlaneCells * 8));
#pragma unroll 8
    for (int k = 0; k < kPlaneF4 / kNormThreads; ++k) {
        const int i = tid + k * kNormThreads;
        const float4 v = __ldg(plane + i);
        out_re[i] = make_float2((v.x - mean[0]) * rstd[0], (v.z - mean[2]) * rstd[2]);
        out_im[i] = make_float2((v.y - mean[1]) * rstd[1], (v.w - mean[3]) * rstd[3]);
    }
}

// One thread per (slot, range, azimuth) position: 16 plane means over elevation, 32 output channels.
__global__ void __launch_bounds__(256)
mnet_kernel(const float* __restrict__ vrdae, const float* __restrict__ weight, const float* __restrict__ bias,
            __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, int n_slots) {
    pdl_wait();                  // programmatic dependent launch: the predecessor's writes are visible from here on (common.cuh)
    pdl_launch_dependents();     // the successor may start its prologue now; it waits the same way before touching memory
    __shared__ float sW[32 * 4 + 32];
    if (threadIdx.x < 160) sW[threadIdx.x] = threadIdx.x < 128 ? __ldg(weight + threadIdx.x) : __ldg(bias + threadIdx.x - 128);
    __syncthreads();
    const size_t gid = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (gid >= (size_t)n_slots * kPlaneCells) return;
    const size_t slot = gid / kPlaneCells;
    const int pos = (int)(gid % kPlaneCells);
    const float4* base = reinterpret_cast<const float4*>(vrdae + slot * (16 * kPlaneCells * 8) + (size_t)pos * 8);
    float m[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float4 a = __ldg(base + (size_t)j * (kPlaneCells * 2));
        const float4 b = __ldg(base + (size_t)j * (kPlaneCells * 2) + 1);
        m[j] = (((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w))) * 0.125f;
    }
    // plane j = chirp f * 2 + (re|im); the reference's .view makes channel 0 = planes 0..7, channel 1 = planes 8..15 (depth = j & 7)
    __nv_bfloat16* oh = out_hi + gid * 32;
    __nv_bfloat16* ol = out_lo ? out_lo + gid * 32 : nullptr;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int o = g * 8 + i;
            const float w00 = sW[o * 4], w01 = sW[o * 4 + 1], w10 = sW[o * 4 + 2], w11 = sW[o * 4 + 3], b = sW[128 + o];
            float best = -INFINITY;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float y = fmaf(w11, m[9 + 2 * t], fmaf(w10, m[8 + 2 * t], fmaf(w01, m[2 * t + 1], w00 * m[2 * t]))) + b;
                best = fmaxf(best, y);
            }
            v[i] = best;
        }
        store8(oh + g * 8, ol ? ol + g * 8 : nullptr, v);
    }
}

// ---- streaming variant: consecutive windows share 7 of 8 frames, and both the standardisation and MNet are
// per-frame operations, so the stream computes them once per frame-sensor instead of once per window slot. ------------------------
// Plane statistics: one CTA per (frame-sensor, kept Doppler row) -> mean and 1/std (unbiased) of the 16 (re|im, elevation) planes.
__global__ void __launch_bounds__(kNormThreads)
plane_stats_kernel(const float4* __restrict__ cube, float* __restrict__ stats /* [n_fs][8][2 mean|rstd][4 quad][4] */) {
    pdl_wait();                  // programmatic dependent launch: the predecessor's writes are visible from here on (common.cuh)
    pdl_launch_dependents();     // the successor may start its prologue now; it waits the same way before touching memory
    __shared__ float scratch[(kNormThreads / 32) * 16];
    const int c = blockIdx.x, fs = blockIdx.y;
    const float4* plane = cube + ((size_t)fs * 16 + 4 + c) * kPlaneF4;
    const int tid = threadIdx.x;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int k = 0; k < kPlaneF4 / kNormThreads; ++k) {
        const float4 v = __ldg(plane + tid + k * kNormThreads);
        s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
    }
    reduce_by_quad(s, scratch);
    float mean[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) mean[k] = s[k] * (1.0f / kPlaneCells);
    float q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int k = 0; k < kPlaneF4 / kNormThreads; ++k) {
        const float4 v = __ldg(plane + tid + k * kNormThreads);
        const float d0 = v.x - mean[0], d1 = v.y - mean[1], d2 = v.z - mean[2], d3 = v.w - mean[3];
        q[0] = fmaf(d0, d0, q[0]); q[1] = fmaf(d1, d1, q[1]); q[2] = fmaf(d2, d2, q[2]); q[3] = fmaf(d3, d3, q[3]);
    }
    reduce_by_quad(q, scratch);
    if (tid < 4) {   // thread t holds the statistics of float4 lane pattern (t & 3): components (re e0, im e0, re e1, im e1), e0 = 2t
        float* dst = stats + ((size_t)(fs * 8 + c) * 2) * 16 + tid * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            dst[k] = mean[k];
            dst[16 + k] = 1.0f / sqrtf(q[k] * (1.0f / (kPlaneCells - 1)));
        }
    }
}

// One thread per (frame-sensor, range, azimuth): standardise with the plane statistics, mean over elevation, MNet.
__global__ void __launch_bounds__(256)
frame_features_kernel(const float4* __restrict__ cube, const float* __restrict__ stats, const float* __restrict__ weight,
                      const float* __restrict__ bias, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
                      int fs0, int n_fs) {
    pdl_wait();                  // programmatic dependent launch: the predecessor's writes are visible from here on (common.cuh)
    pdl_launch_dependents();     // the successor may start its prologue now; it waits the same way before touching memory
    __shared__ float sW[32 * 4 + 32];
    __shared__ float sStat[8 * 32];
    if (threadIdx.x < 160) sW[threadIdx.x] = threadIdx.x < 128 ? __ldg(weight + threadIdx.x) : __ldg(bias + threadIdx.x - 128);
    const size_t gid = (size_t)blockIdx.x * 256 + threadIdx.x;       // 4096 positions per frame-sensor, 256 | 4096 -> one fs per CTA
    const int fs = fs0 + (int)(gid / kPlaneCells);
    const int pos = (int)(gid % kPlaneCells);
    sStat[threadIdx.x] = __ldg(stats + (size_t)fs * 256 + threadIdx.x);
    __syncthreads();
    float m[16];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4* src = cube + ((size_t)fs * 16 + 4 + c) * kPlaneF4 + (size_t)pos * 4;
        const float* mean = sStat + c * 32;
        const float* rstd = mean + 16;
        float re = 0.f, im = 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float4 v = __ldg(src + t);
            re += (v.x - mean[t * 4]) * rstd[t * 4] + (v.z - mean[t * 4 + 2]) * rstd[t * 4 + 2];
            im += (v.y - mean[t * 4 + 1]) * rstd[t * 4 + 1] + (v.w - mean[t * 4 + 3]) * rstd[t * 4 + 3];
        }
        m[2 * c] = re * 0.125f;
        m[2 * c + 1] = im * 0.125f;
    }
    const size_t o = ((size_t)(fs - fs0) * kPlaneCells + pos) * 32;
    __nv_bfloat16* oh = out_hi + o;
    __nv_bfloat16* ol = out_lo ? out_lo + o : nullptr;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int oc = g * 8 + i;
            const float w00 = sW[oc * 4], w01 = sW[oc * 4 + 1], w10 = sW[oc * 4 + 2], w11 = sW[oc * 4 + 3], b = sW[128 + oc];
            float best = -INFINITY;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float y = fmaf(w11, m[9 + 2 * t], fmaf(w10, m[8 + 2 * t], fmaf(w01, m[2 * t + 1], w00 * m[2 * t]))) + b;
                best = fmaxf(best, y);
            }
            v[i] = best;
        }
        store8(oh + g * 8, ol ? ol + g * 8 : nullptr, v);
    }
}

struct ResampleParams {
    int n, di, hi, wi, dout, ho, wo, c8;      // c8 = channels / 8
    int in_ld, in_off, out_ld, out_off;
    float sd, sh, sw;                         // align_corners scales (in-1)/(out-1)
};

__global__ void __launch_bounds__(256)
resample_kernel(const __nv_bfloat16* __restrict__ in_hi, const __nv_bfloat16* __restrict__ in_lo,
                __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, const ResampleParams p) {
    pdl_wait();                  // programmatic dependent launch: the predecessor's writes are visible from here on (common.cuh)
    pdl_launch_dependents();     // the successor may start its prologue now; it waits the same way before touching memory
    const size_t total = (size_t)p.n * p.dout * p.ho * p.wo * p.c8;
    for (size_t gid = (size_t)blockIdx.x * 256 + threadIdx.x; gid < total; gid += (size_t)gridDim.x * 256) {
        size_t t = gid;
        const int cg = (int)(t % p.c8); t /= p.c8;
        const int ow = (int)(t % p.wo); t /= p.wo;
        const int oh = (int)(t % p.ho); t /= p.ho;
        const int od = (int)(t % p.dout);
        const int n = (int)(t / p.dout);
        const float fd = p.sd * od, fh = p.sh * oh, fw = p.sw * ow;
        const int d0 = (int)fd, h0 = (int)fh, w0 = (int)fw;
        const int d1 = d0 + (d0 < p.di - 1), h1 = h0 + (h0 < p.hi - 1), w1 = w0 + (w0 < p.wi - 1);
        const float ld1 = fd - d0, lh1 = fh - h0, lw1 = fw - w0;
        const float ld0 = 1.f - ld1, lh0 = 1.f - lh1, lw0 = 1.f - lw1;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        const int ds[2] = {d0, d1}, hs[2] = {h0, h1}, ws[2] = {w0, w1};
        const float wd[2] = {ld0, ld1}, wh[2] = {lh0, lh1}, ww[2] = {lw0, lw1};
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            if (a == 1 && p.di == 1) break;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const float wgt = (p.di == 1 ? 1.f : wd[a]) * wh[b] * ww[c];
                    const size_t off = ((((size_t)n * p.di + ds[a]) * p.hi + hs[b]) * p.wi + ws[c]) * p.in_ld + p.in_off + cg * 8;
                    float v[8];
                    load8(in_hi + off, in_lo ? in_lo + off : nullptr, v);
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i] = fmaf(wgt, v[i], acc[i]);
                }
            }
        }
        const size_t ooff = ((((size_t)n * p.dout + od) * p.ho + oh) * p.wo + ow) * p.out_ld + p.out_off + cg * 8;
        store8(out_hi + ooff, out_lo ? out_lo + ooff : nullptr, acc);
    }
}

// One CTA (256 threads) per row; cols in {256, 512, ..., 4096} (multiple of 4, <= 4096).
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ logits, __nv_bfloat16* __restrict__ p_hi, __nv_bfloat16* __restrict__ p_lo, int cols) {
    pdl_wait();                  // programmatic dependent launch: the predecessor's writes are visible from here on (common.cuh)
    pdl_launch_dependents();     // the successor may start its prologue now; it waits the same way before touching memory
    __shared__ float red[8];
    const size_t row = blockIdx.x;
    const float4* src = reinterpret_cast<const float4*>(logits + row * cols);
    const int nvec = cols >> 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float4 v[4];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = tid + k * 256;
        if (i < nvec) {
            v[k] = __ldg(src + i);
            mx = fmaxf(mx, fmaxf(fmaxf(v[k].x, v[k].y), fmaxf(v[k].z, v[k].w)));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = tid + k * 256;
        if (i < nvec) {
            v[k].x = expf(v[k].x - mx); v[k].y = expf(v[k].y - mx); v[k].z = expf(v[k].z - mx); v[k].w = expf(v[k].w - mx);
            sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w];
    const float inv = 1.0f / sum;
    uint2* dh = reinterpret_cast<uint2*>(p_hi + row * cols);
    uint2* dl = p_lo ? reinterpret_cast<uint2*>(p_lo + row * cols) : nullptr;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = tid + k * 256;
        if (i < nvec) {
            uint32_t h0, l0, h1, l1;
            split2(v[k].x * inv, v[k].y * inv, h0, l0);
            split2(v[k].z * inv, v[k].w * inv, h1, l1);
            dh[i] = make_uint2(h0, h1);
            if (dl) dl[i] = make_uint2(l0, l1);
        }
    }
}

// in [n][s][in_ld] (channels in_off .. in_off + c) -> out [n][c][s]; 32x32 tiles, block (32, 8).
__global__ void __launch_bounds__(256)
transpose_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int s, int c, int in_ld, int in_off) {
    pdl_wait();                  // programmatic dependent launch: the predecessor's writes are visible from here on (common.cuh)
    pdl_launch_dependents();     // the successor may start its prologue now; it waits the same way before touching memory
    __shared__ uint16_t tile[32][34];
    const int n = blockIdx.z;
    const int s0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
    for (int j = ty; j < 32; j += 8) tile[j][tx] = in[((size_t)n * s + s0 + j) * in_ld + in_off + c0 + tx];
    __syncthreads();
#pragma unroll
    for (int j = ty; j < 32; j += 8) out[((size_t)n * c + c0 + j) * s + s0 + tx] = tile[tx][j];
}

static int check_sm100() {
    return device_check_sm100();      // cached per device (capi.cu)
}

int fast_transpose_dense(const void* in_hi, const void* in_lo, int n, int s, int c, int in_ld, int in_ch_off, void* out_hi, void* out_lo,
                         cudaStream_t stream);
int fast_to_kmajor(const void* src_hi, const void* src_lo, int n, int d, int h, int w, int ld, int ch_off, int c, void* dst_hi, void* dst_lo,
                   int dp, int hp, int wp, int pd, int ph, int pw, int shift0, int n_copies, long long copy_stride, long long ppad,
                   cudaStream_t stream);

static inline int launch_status(int n = 1) {
    note_launches(n);
    return cudaGetLastError() == cudaSuccess ? HUPR_OK : HUPR_ERR_CUDA;
}

}  // namespace hupr

using namespace hupr;

extern "C" int hupr_window_normalize(const void* cube, const int32_t* slot_fs, int n_slots, float* vrdae, void* stream) {
    if (n_slots < 0) return HUPR_ERR_BAD_ARG;
    if (n_slots == 0) return HUPR_OK;
    if (!cube || !slot_fs || !vrdae) return HUPR_ERR_BAD_ARG;
    if (((uintptr_t)cube | (uintptr_t)vrdae) & 15) return HUPR_ERR_ALIGNMENT;
    int rc = check_sm100();
    if (rc != HUPR_OK) return rc;
    launch_k(window_normalize_kernel, dim3(dim3(8, n_slots)), dim3(kNormThreads), (size_t)(0), (cudaStream_t)stream, 
        static_cast<const float4*>(cube), slot_fs, vrdae);
    return launch_status();
}

extern "C" size_t hupr_frame_features_workspace_bytes(int n_frame_sensors) {
    return n_frame_sensors > 0 ? (size_t)n_frame_sensors * 256 * sizeof(float) : 0;
}

extern "C" int hupr_plane_stats(const void* cube, int n_frame_sensors, void* workspace, size_t ws_bytes, void* stream) {
    if (n_frame_sensors < 0) return HUPR_ERR_BAD_ARG;
    if (n_frame_sensors == 0) return HUPR_OK;
    if (!cube || !workspace) return HUPR_ERR_BAD_ARG;
    if (ws_bytes < hupr_frame_features_workspace_bytes(n_frame_sensors)) return HUPR_ERR_WORKSPACE;
    if (((uintptr_t)cube | (uintptr_t)workspace) & 15) return HUPR_ERR_ALIGNMENT;
    int rc = check_sm100();
    if (rc != HUPR_OK) return rc;
    launch_k(plane_stats_kernel, dim3(dim3(8, n_frame_sensors)), dim3(kNormThreads), (size_t)(0), (cudaStream_t)stream, static_cast<const float4*>(cube),
                                                                                          static_cast<float*>(workspace));
    return launch_status();
}

extern "C" int hupr_frame_features(const void* cube, const void* stats, int first_frame_sensor, int n_frame_sensors, const float* weight,
                                   const float* bias, void* out_hi, void* out_lo, void* stream) {
    if (n_frame_sensors < 0 || first_frame_sensor < 0) return HUPR_ERR_BAD_ARG;
    if (n_frame_sensors == 0) return HUPR_OK;
    if (!cube || !stats || !weight || !bias || !out_hi) return HUPR_ERR_BAD_ARG;
    if (((uintptr_t)cube | (uintptr_t)out_hi | (uintptr_t)out_lo) & 15) return HUPR_ERR_ALIGNMENT;
    int rc = check_sm100();
    if (rc != HUPR_OK) return rc;
    launch_k(frame_features_kernel, dim3(n_frame_sensors * (kPlaneCells / 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, 
        static_cast<const float4*>(cube), static_cast<const float*>(stats), weight, bias, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo,
        first_frame_sensor, n_frame_sensors);
    return launch_status();
}

extern "C" int hupr_mnet_fwd(const float* vrdae, const float* weight, const float* bias, void* out_hi, void* out_lo,
                             int n_slots, void* stream) {
    if (n_slots < 0) return HUPR_ERR_BAD_ARG;
    if (n_slots == 0) return HUPR_OK;
    if (!vrdae || !weight || !bias || !out_hi) return HUPR_ERR_BAD_ARG;
    if (((uintptr_t)vrdae | (uintptr_t)out_hi | (uintptr_t)out_lo) & 15) return HUPR_ERR_ALIGNMENT;
    int rc = check_sm100();
    if (rc != HUPR_OK) return rc;
    const size_t total = (size_t)n_slots * kPlaneCells;
    launch_k(mnet_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, 
        vrdae, weight, bias, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, n_slots);
    return launch_status();
}

extern "C" int hupr_resample_linear(const void* in_hi, const void* in_lo, int n, int di, int hi, int wi, int c, int in_ld, int in_ch_off,
                                    void* out_hi, void* out_lo, int dout, int ho, int wo, int out_ld, int out_ch_off, void* stream) {
    if (!in_hi || !out_hi || n <= 0 || di <= 0 || hi <= 0 || wi <= 0 || dout <= 0 || ho <= 0 || wo <= 0 || c <= 0) return HUPR_ERR_BAD_ARG;
    if (c % 8 || in_ld % 8 || out_ld % 8 || in_ch_off % 8 || out_ch_off % 8 || in_ch_off + c > in_ld || out_ch_off + c > out_ld)
        return HUPR_ERR_BAD_ARG;
    if ((in_lo == nullptr) != (out_lo == nullptr)) return HUPR_ERR_BAD_ARG;
    if (((uintptr_t)in_hi | (uintptr_t)in_lo | (uintptr_t)out_hi | (uintptr_t)out_lo) & 15) return HUPR_ERR_ALIGNMENT;
    int rc = check_sm100();
    if (rc != HUPR_OK) return rc;
    ResampleParams p;
    p.n = n; p.di = di; p.hi = hi; p.wi = wi; p.dout = dout; p.ho = ho; p.wo = wo; p.c8 = c / 8;
    p.in_ld = in_ld; p.in_off = in_ch_off; p.out_ld = out_ld; p.out_off = out_ch_off;
    p.sd = dout > 1 ? (float)(di - 1) / (float)(dout - 1) : 0.f;
    p.sh = ho > 1 ? (float)(hi - 1) / (float)(ho - 1) : 0.f;
    p.sw = wo > 1 ? (float)(wi - 1) / (float)(wo - 1) : 0.f;
    const size_t total = (size_t)n * dout * ho * wo * p.c8;
    size_t blocks = (total + 255) / 256;
    if (blocks > 148 * 64) blocks = 148 * 64;
    launch_k(resample_kernel, dim3((unsigned)blocks), dim3(256), (size_t)(0), (cudaStream_t)stream, (const __nv_bfloat16*)in_hi, (const __nv_bfloat16*)in_lo,
                                                                       (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, p);
    return launch_status();
}

extern "C" int hupr_softmax_rows(const float* logits, void* p_hi, void* p_lo, long long rows, int cols, void* stream) {
    if (rows < 0 || rows > 2147483647LL) return HUPR_ERR_BAD_ARG;
    if (rows == 0) return HUPR_OK;
    if (!logits || !p_hi || cols <= 0 || cols % 4 || cols > 4096) return HUPR_ERR_BAD_ARG;
    if (((uintptr_t)logits | (uintptr_t)p_hi | (uintptr_t)p_lo) & 15) return HUPR_ERR_ALIGNMENT;
    int rc = check_sm100();
    if (rc != HUPR_OK) return rc;
    launch_k(softmax_rows_kernel, dim3((unsigned)rows), dim3(256), (size_t)(0), (cudaStream_t)stream, logits, (__nv_bfloat16*)p_hi, (__nv_bfloat16*)p_lo, cols);
    return launch_status();
}

extern "C" int hupr_transpose_split(const void* in_hi, const void* in_lo, int n, int s, int c, int in_ld, int in_ch_off,
                                    void* out_hi, void* out_lo, void* stream) {
    if (!in_hi || !out_hi || n <= 0 || s <= 0 || c <= 0 || s % 32 || c % 32 || in_ch_off < 0 || in_ch_off + c > in_ld) return HUPR_ERR_BAD_ARG;
    if ((in_lo == nullptr) != (out_lo == nullptr)) return HUPR_ERR_BAD_ARG;
    if (n > 65535 || c / 32 > 65535) return HUPR_ERR_BAD_ARG;
    int rc = check_sm100();
    if (rc != HUPR_OK) return rc;
    if (s % 64 == 0 && c % 64 == 0 && in_ld % 8 == 0 && in_ch_off % 8 == 0 &&
        !(((uintptr_t)in_hi | (uintptr_t)in_lo) & 15) && !(((uintptr_t)out_hi | (uintptr_t)out_lo) & 3))
        return fast_transpose_dense(in_hi, in_lo, n, s, c, in_ld, in_ch_off, out_hi, out_lo, (cudaStream_t)stream);      // transpose.cu
    const dim3 grid(s / 32, c / 32, n), block(32, 8);
    launch_k(transpose_kernel, dim3(grid), dim3(block), (size_t)(0), (cudaStream_t)stream, (const uint16_t*)in_hi, (uint16_t*)out_hi, s, c, in_ld, in_ch_off);
    if (in_lo) launch_k(transpose_kernel, dim3(grid), dim3(block), (size_t)(0), (cudaStream_t)stream, (const uint16_t*)in_lo, (uint16_t*)out_lo, s, c, in_ld, in_ch_off);
    return launch_status(in_lo ? 2 : 1);
}

// ================================================================================================================================
// Position-major ("K-major") copies for the weight-gradient GEMMs.
//   dW[tap][co][ci] = sum_pos dY[pos][co] * X[pos + off(tap)][ci]
// contracts over POSITIONS, so both operands are re-laid out as [channel][P] with P a zero-padded linear position index
//   P(n, d, h, w) = ((n*Dp + d + pd)*Hp + h + ph)*Wp + w + pw
// in which a filter tap is a constant shift `off` (hupr_conv_desc.w_k_off) and the convolution's zero padding is literal zeros.
// TMA needs 16-byte aligned box starts, so the +-1 shifts along W cannot be coordinates: the X operand is stored once per kw tap,
// pre-shifted by `shift` = kw - pw elements, and Wp is a multiple of 8 so that the kd / kh parts of `off` stay aligned.
// ================================================================================================================================
namespace hupr {

struct KMajorParams {
    int n, d, h, w, ld, ch_off;
    int dp, hp, wp, pd, ph, pw, shift;
    long long ppad;
};

__global__ void __launch_bounds__(256)
to_kmajor_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, const KMajorParams p) {
    __shared__ uint16_t tile[32][34];
    const long long pos0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
    for (int j = ty; j < 32; j += 8) tile[j][tx] = src[(size_t)(pos0 + j) * p.ld + p.ch_off + c0 + tx];
    __syncthreads();
    long long pos = pos0 + tx;
    const int w = (int)(pos % p.w); pos /= p.w;
    const int h = (int)(pos % p.h); pos /= p.h;
    const int d = (int)(pos % p.d);
    const int n = (int)(pos / p.d);
    const long long P = (((long long)n * p.dp + d + p.pd) * p.hp + h + p.ph) * p.wp + w + p.pw - p.shift;
#pragma unroll
    for (int j = ty; j < 32; j += 8) dst[(size_t)(c0 + j) * p.ppad + P] = tile[tx][j];
}

}  // namespace hupr

extern "C" int hupr_to_kmajor(const void* src_hi, const void* src_lo, int n, int d, int h, int w, int ld, int ch_off, int c,
                              void* dst_hi, void* dst_lo, int dp, int hp, int wp, int pd, int ph, int pw, int shift, long long ppad,
                              void* stream) {
    if (!src_hi || !dst_hi || n <= 0 || d <= 0 || h <= 0 || w <= 0 || c <= 0 || c % 32 || ch_off < 0 || ch_off + c > ld) return HUPR_ERR_BAD_ARG;
    if ((src_lo == nullptr) != (dst_lo == nullptr)) return HUPR_ERR_BAD_ARG;
    const long long positions = (long long)n * d * h * w;
    if (positions % 32 || d + pd > dp || h + ph > hp || w + pw > wp || pd < 0 || ph < 0 || pw < 0) return HUPR_ERR_BAD_ARG;
    if ((long long)(pd * hp + ph) * wp + pw - shift < 0 || shift < -pw - 1) return HUPR_ERR_BAD_ARG;       // first / last interior cell stay inside dst
    if (ppad < (long long)n * dp * hp * wp || positions / 32 > 2147483647LL || c / 32 > 65535) return HUPR_ERR_BAD_ARG;
    int rc = check_sm100();
    if (rc != HUPR_OK) return rc;
    if (positions % 64 == 0 && c % 64 == 0 && w % 2 == 0 && ld % 8 == 0 && ch_off % 8 == 0 && ppad % 2 == 0 &&
        !(((uintptr_t)src_hi | (uintptr_t)src_lo) & 15) && !(((uintptr_t)dst_hi | (uintptr_t)dst_lo) & 3))
        return fast_to_kmajor(src_hi, src_lo, n, d, h, w, ld, ch_off, c, dst_hi, dst_lo, dp, hp, wp, pd, ph, pw, shift, 1, 0, ppad,
                              (cudaStream_t)stream);                                                                     // transpose.cu
    KMajorParams p;
    p.n = n; p.d = d; p.h = h; p.w = w; p.ld = ld; p.ch_off = ch_off;
    p.dp = dp; p.hp = hp; p.wp = wp; p.pd = pd; p.ph = ph; p.pw = pw; p.shift = shift; p.ppad = ppad;
    const dim3 grid((unsigned)(positions / 32), c / 32), block(32, 8);
    to_kmajor_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const uint16_t*)src_hi, (uint16_t*)dst_hi, p);
    if (src_lo) to_kmajor_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const uint16_t*)src_lo, (uint16_t*)dst_lo, p);
    return launch_status(src_lo ? 2 : 1);
}
