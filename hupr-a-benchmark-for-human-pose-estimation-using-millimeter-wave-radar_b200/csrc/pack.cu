// Parameter-layout kernels of the training step (sm_100a; tiny, HBM/latency-bound).
//
// The optimiser (hupr_adam_step) updates ONE flat fp32 parameter buffer in torch's layouts ([cout][cin][kd][kh][kw], PReLU scalars,
// GCN [1024][1024] / [1024][14]); the tensor-core kernels consume hi/lo bf16 operand layouts ([tap][cout][cin] forward,
// [flipped tap][cin][cout] data-gradient), and hupr_conv_wgrad produces [tap][cin][cout] fp32.  These kernels move between the two
// worlds inside the captured step, so that no framework element-wise kernel runs between the optimiser and the next forward
// (what torch.optim.Adam + autograd do implicitly for /root/reference/tools/run.py:76-79):
//   hupr_pack_conv_weights   fp32 [cout][cin][taps] -> forward AND data-gradient operand layouts (hi/lo bf16), one read
//   hupr_unpack_wgrad        fp32 [taps][cin_pad][cout_total] accumulator -> fp32 [cout][cin][taps] gradient (torch layout)
//   hupr_reduce_f64          float dst[g] = sum_i double src[g*n + i]        (double partial sums -> fp32 gradients / casts)
//   hupr_broadcast_f32       dst[0..n) = *src                                (PReLU scalar -> per-channel slope array)
//   hupr_gcn_bias_rows       bias fp32 [1024][14] -> hi/lo rows [(b,j)][1024] (residual operand of the transposed-layout GCN GEMMs)
//   hupr_bump_i32            ++*counter                                      (device-side Adam step count)
//   hupr_memset_zero         cudaMemsetAsync (a memset node when captured)
// and, for the inference path, hupr_quantize_planes: bf16 hi/lo -> the fp16 + e4m3 operand planes of the two-unit 3-tap convolution
// (hupr_conv_desc.nprod == 2); HBM-bound, 4 B read + 4 B written per element.
#include "common.cuh"
#include "split.cuh"

namespace hupr {

constexpr int PK_T = 16;     // channels per tile side

// Job table of the batched entry points (hupr_pack_conv_weights_multi / hupr_unpack_wgrad_multi): the ~85 parameter tensors of the model
// are (un)packed by ONE launch each; `block_job[b]` names the job of CTA b and `block_begin` its first CTA.
struct PackJob {                 // mirrors hupr_pack_job (include/hupr_b200.h)
    const float* w;              // pack: fp32 source [cout][cin][taps]; unpack: fp32 accumulator [taps][cin_pad][cout_total]
    void* a_hi; void* a_lo;      // pack: forward operand planes; unpack: a_hi = fp32 destination [cout][cin][taps], a_lo unused
    void* b_hi; void* b_lo;      // pack: data-gradient operand planes (may be null); unpack: unused
    int cout, cin, taps;
    int cout_total, cin_pad, cout_off;
    int blocks_x;                // ceil(cout / 16)
    int block_begin;             // first CTA of this job in the batched grid
};

// One CTA: a [16 cout] x [16 cin] x [taps] brick.  Source rows (ci, tap) of one cout are contiguous, so the load is coalesced; both
// stores write 32-byte segments (16 bf16 along cin for the forward layout, 16 bf16 along cout for the data-gradient layout).
__device__ __forceinline__ void
pack_conv_brick(const float* __restrict__ w, int cout, int cin, int taps, __nv_bfloat16* __restrict__ f_hi, __nv_bfloat16* __restrict__ f_lo,
                int cout_total, int cin_pad, int cout_off, __nv_bfloat16* __restrict__ d_hi, __nv_bfloat16* __restrict__ d_lo, int bx, int by) {
    extern __shared__ float brick[];                      // [16 co][16 ci][taps]
    const int co0 = bx * PK_T, ci0 = by * PK_T;
    const int row = PK_T * taps;                          // floats of one cout row of the brick
    for (int i = threadIdx.x; i < PK_T * row; i += 256) {
        const int co_l = i / row, rem = i - co_l * row;
        const int ci_l = rem / taps;
        float v = 0.f;
        if (co0 + co_l < cout && ci0 + ci_l < cin) v = __ldg(w + ((size_t)(co0 + co_l) * cin + ci0) * taps + rem);
        brick[i] = v;
    }
    __syncthreads();
    // forward layout [tap][cout_total][cin_pad]: pairs along cin
    for (int i = threadIdx.x; i < taps * PK_T * (PK_T / 2); i += 256) {
        const int cp = i % (PK_T / 2), co_l = (i / (PK_T / 2)) % PK_T, t = i / (PK_T * PK_T / 2);
        const int ci = ci0 + 2 * cp, co = co0 + co_l;
        if (co >= cout || ci >= cin) continue;
        uint32_t hi, lo;
        split2(brick[(co_l * PK_T + 2 * cp) * taps + t], brick[(co_l * PK_T + 2 * cp + 1) * taps + t], hi, lo);
        const size_t o = ((size_t)t * cout_total + cout_off + co) * cin_pad + ci;
        *reinterpret_cast<uint32_t*>(f_hi + o) = hi;
        if (f_lo) *reinterpret_cast<uint32_t*>(f_lo + o) = lo;
    }
    if (!d_hi) return;
    // data-gradient layout [taps - 1 - tap][cin_pad][cout_total]: pairs along cout
    for (int i = threadIdx.x; i < taps * PK_T * (PK_T / 2); i += 256) {
        const int cp = i % (PK_T / 2), ci_l = (i / (PK_T / 2)) % PK_T, t = i / (PK_T * PK_T / 2);
        const int ci = ci0 + ci_l, co = co0 + 2 * cp;
        if (co >= cout || ci >= cin) continue;
        uint32_t hi, lo;
        split2(brick[((2 * cp) * PK_T + ci_l) * taps + t], brick[((2 * cp + 1) * PK_T + ci_l) * taps + t], hi, lo);
        const size_t o = ((size_t)(taps - 1 - t) * cin_pad + ci) * cout_total + cout_off + co;
        *reinterpret_cast<uint32_t*>(d_hi + o) = hi;
        if (d_lo) *reinterpret_cast<uint32_t*>(d_lo + o) = lo;
    }
}

__global__ void __launch_bounds__(256)
pack_conv_kernel(const float* __restrict__ w, int cout, int cin, int taps, __nv_bfloat16* __restrict__ f_hi, __nv_bfloat16* __restrict__ f_lo,
                 int cout_total, int cin_pad, int cout_off, __nv_bfloat16* __restrict__ d_hi, __nv_bfloat16* __restrict__ d_lo) {
    pack_conv_brick(w, cout, cin, taps, f_hi, f_lo, cout_total, cin_pad, cout_off, d_hi, d_lo, blockIdx.x, blockIdx.y);
}

// acc fp32 [taps][cin_pad][cout_total] -> dst fp32 [cout][cin][taps] for the output channels [cout_off, cout_off + cout).
__device__ __forceinline__ void
unpack_wgrad_brick(const float* __restrict__ acc, int taps, int cin_pad, int cout_total, int cout_off, float* __restrict__ dst, int cout, int cin,
                   int bx, int by) {
    extern __shared__ float brick[];                      // [16 co][16 ci][taps]
    const int co0 = bx * PK_T, ci0 = by * PK_T;
    for (int i = threadIdx.x; i < taps * PK_T * PK_T; i += 256) {
        const int co_l = i % PK_T, ci_l = (i / PK_T) % PK_T, t = i / (PK_T * PK_T);
        float v = 0.f;
        if (co0 + co_l < cout && ci0 + ci_l < cin) v = __ldg(acc + ((size_t)t * cin_pad + ci0 + ci_l) * cout_total + cout_off + co0 + co_l);
        brick[(co_l * PK_T + ci_l) * taps + t] = v;
    }
    __syncthreads();
    const int row = PK_T * taps;
    for (int i = threadIdx.x; i < PK_T * row; i += 256) {
        const int co_l = i / row, rem = i - co_l * row;
        const int ci_l = rem / taps;
        if (co0 + co_l < cout && ci0 + ci_l < cin) dst[((size_t)(co0 + co_l) * cin + ci0) * taps + rem] = brick[i];
    }
}

__global__ void __launch_bounds__(256)
unpack_wgrad_kernel(const float* __restrict__ acc, int taps, int cin_pad, int cout_total, int cout_off, float* __restrict__ dst, int cout, int cin) {
    unpack_wgrad_brick(acc, taps, cin_pad, cout_total, cout_off, dst, cout, cin, blockIdx.x, blockIdx.y);
}

__global__ void __launch_bounds__(256)
pack_conv_multi_kernel(const PackJob* __restrict__ jobs, const int* __restrict__ block_job) {
    const PackJob j = jobs[__ldg(block_job + blockIdx.x)];
    const int local = blockIdx.x - j.block_begin;
    pack_conv_brick(j.w, j.cout, j.cin, j.taps, (__nv_bfloat16*)j.a_hi, (__nv_bfloat16*)j.a_lo, j.cout_total, j.cin_pad, j.cout_off,
                    (__nv_bfloat16*)j.b_hi, (__nv_bfloat16*)j.b_lo, local % j.blocks_x, local / j.blocks_x);
}

__global__ void __launch_bounds__(256)
unpack_wgrad_multi_kernel(const PackJob* __restrict__ jobs, const int* __restrict__ block_job) {
    const PackJob j = jobs[__ldg(block_job + blockIdx.x)];
    const int local = blockIdx.x - j.block_begin;
    unpack_wgrad_brick(j.w, j.taps, j.cin_pad, j.cout_total, j.cout_off, (float*)j.a_hi, j.cout, j.cin, local % j.blocks_x, local / j.blocks_x);
}

__global__ void __launch_bounds__(128)
reduce_f64_kernel(const double* __restrict__ src, int groups, int n, float* __restrict__ dst) {
    const int g = blockIdx.x * 128 + threadIdx.x;
    if (g >= groups) return;
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += src[(size_t)g * n + i];
    dst[g] = (float)s;
}

// thread = 8 consecutive channels of one row; two independent items per iteration so that four 16-byte loads are in flight
__global__ void __launch_bounds__(256)
quantize_planes_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, long long items, int ld, int ch_off, int groups,
                       __half* __restrict__ q16, uint8_t* __restrict__ q8, int is_weight, int* __restrict__ sat) {
    const QuantScales q = quant_scales(is_weight != 0);
    const long long stride = (long long)gridDim.x * 256;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < items; i += stride) {
        const long long row = i / groups;
        const size_t at = (size_t)row * ld + ch_off + (size_t)(i - row * groups) * 8;
        float v[8];
        load8(hi + at, lo ? lo + at : nullptr, v);
        uint4 h;
        uint2 a8, l8;
        quant_report(sat, quant8(v, q, h, a8, l8));
        *reinterpret_cast<uint4*>(q16 + at) = h;
        // e4m3 plane: [row][ld / 32][2][32] — the 32 values of a channel block, then its 32 residuals
        const int c = ch_off + (int)(i - row * groups) * 8;
        uint8_t* q = q8 + (size_t)row * (2 * ld) + (size_t)(c >> 5) * 64 + (c & 31);
        *reinterpret_cast<uint2*>(q) = a8;
        *reinterpret_cast<uint2*>(q + 32) = l8;
    }
}

__global__ void __launch_bounds__(256)
broadcast_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) dst[i] = __ldg(src);
}

// rows[(b, j)][q] = bias[q][j] for b < batch, zero rows beyond batch*14; thread = (row, pair of q)
__global__ void __launch_bounds__(256)
gcn_bias_rows_kernel(const float* __restrict__ bias, int batch, int rows, __nv_bfloat16* __restrict__ r_hi, __nv_bfloat16* __restrict__ r_lo) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long long)rows * 512) return;
    const int r = (int)(i / 512), q = (int)(i % 512) * 2;
    float a = 0.f, b = 0.f;
    if (r < batch * 14) {
        const int j = r % 14;
        a = __ldg(bias + (size_t)q * 14 + j);
        b = __ldg(bias + (size_t)(q + 1) * 14 + j);
    }
    uint32_t hi, lo;
    split2(a, b, hi, lo);
    *reinterpret_cast<uint32_t*>(r_hi + (size_t)r * 1024 + q) = hi;
    if (r_lo) *reinterpret_cast<uint32_t*>(r_lo + (size_t)r * 1024 + q) = lo;
}

__global__ void bump_i32_kernel(int* p) { *p += 1; }

static inline int pack_status() {
    note_launches(1);
    return cudaGetLastError() == cudaSuccess ? HUPR_OK : HUPR_ERR_CUDA;
}

}  // namespace hupr

extern "C" int hupr_pack_conv_weights(const float* w, int cout, int cin, int taps, void* fwd_hi, void* fwd_lo, int cout_total, int cin_pad,
                                      int cout_off, void* dgrad_hi, void* dgrad_lo, void* stream) {
    using namespace hupr;
    if (!w || !fwd_hi || cout <= 0 || cin <= 0 || taps <= 0 || taps > 32) return HUPR_ERR_BAD_ARG;
    if ((cout & 1) || (cin & 1) || (cout_off & 1) || (cin_pad & 1) || (cout_total & 1)) return HUPR_ERR_BAD_ARG;       // bf16 pairs
    if (cout_off < 0 || cout_off + cout > cout_total || cin > cin_pad) return HUPR_ERR_BAD_ARG;
    if (dgrad_lo && !dgrad_hi) return HUPR_ERR_BAD_ARG;
    if (((uintptr_t)fwd_hi | (uintptr_t)fwd_lo | (uintptr_t)dgrad_hi | (uintptr_t)dgrad_lo) & 3) return HUPR_ERR_ALIGNMENT;
    if (int rc = device_check_sm100()) return rc;
    const dim3 grid((cout + PK_T - 1) / PK_T, (cin + PK_T - 1) / PK_T);
    const size_t smem = (size_t)PK_T * PK_T * taps * sizeof(float);
    pack_conv_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(w, cout, cin, taps, (__nv_bfloat16*)fwd_hi, (__nv_bfloat16*)fwd_lo, cout_total,
                                                              cin_pad, cout_off, (__nv_bfloat16*)dgrad_hi, (__nv_bfloat16*)dgrad_lo);
    return pack_status();
}

extern "C" int hupr_unpack_wgrad(const float* acc, int taps, int cin_pad, int cout_total, int cout_off, float* dst, int cout, int cin,
                                 void* stream) {
    using namespace hupr;
    if (!acc || !dst || cout <= 0 || cin <= 0 || taps <= 0 || taps > 32) return HUPR_ERR_BAD_ARG;
    if (cout_off < 0 || cout_off + cout > cout_total || cin > cin_pad) return HUPR_ERR_BAD_ARG;
    if (int rc = device_check_sm100()) return rc;
    const dim3 grid((cout + PK_T - 1) / PK_T, (cin + PK_T - 1) / PK_T);
    const size_t smem = (size_t)PK_T * PK_T * taps * sizeof(float);
    unpack_wgrad_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(acc, taps, cin_pad, cout_total, cout_off, dst, cout, cin);
    return pack_status();
}

// Batched forms: jobs / block_job are DEVICE arrays built once by the caller (hupr_pack_job: include/hupr_b200.h); every job must satisfy
// the single-job entry point's constraints (validated by the caller when the table is built: hupr_b200/training.py).
extern "C" int hupr_pack_conv_weights_multi(const hupr_pack_job* jobs, const int* block_job, int total_blocks, int max_taps, void* stream) {
    using namespace hupr;
    static_assert(sizeof(PackJob) == sizeof(hupr_pack_job), "PackJob mirrors hupr_pack_job");
    if (total_blocks < 0 || max_taps <= 0 || max_taps > 32) return HUPR_ERR_BAD_ARG;
    if (total_blocks == 0) return HUPR_OK;
    if (!jobs || !block_job) return HUPR_ERR_BAD_ARG;
    if (int rc = device_check_sm100()) return rc;
    const size_t smem = (size_t)PK_T * PK_T * max_taps * sizeof(float);
    pack_conv_multi_kernel<<<total_blocks, 256, smem, (cudaStream_t)stream>>>(reinterpret_cast<const PackJob*>(jobs), block_job);
    return pack_status();
}

extern "C" int hupr_unpack_wgrad_multi(const hupr_pack_job* jobs, const int* block_job, int total_blocks, int max_taps, void* stream) {
    using namespace hupr;
    if (total_blocks < 0 || max_taps <= 0 || max_taps > 32) return HUPR_ERR_BAD_ARG;
    if (total_blocks == 0) return HUPR_OK;
    if (!jobs || !block_job) return HUPR_ERR_BAD_ARG;
    if (int rc = device_check_sm100()) return rc;
    const size_t smem = (size_t)PK_T * PK_T * max_taps * sizeof(float);
    unpack_wgrad_multi_kernel<<<total_blocks, 256, smem, (cudaStream_t)stream>>>(reinterpret_cast<const PackJob*>(jobs), block_job);
    return pack_status();
}

extern "C" int hupr_reduce_f64(const double* src, int groups, int n, float* dst, void* stream) {
    using namespace hupr;
    if (groups < 0 || n <= 0) return HUPR_ERR_BAD_ARG;
    if (groups == 0) return HUPR_OK;
    if (!src || !dst) return HUPR_ERR_BAD_ARG;
    if (int rc = device_check_sm100()) return rc;
    reduce_f64_kernel<<<(groups + 127) / 128, 128, 0, (cudaStream_t)stream>>>(src, groups, n, dst);
    return pack_status();
}

extern "C" int hupr_quantize_planes(const void* hi, const void* lo, long long rows, int ld, int ch_off, int ch, void* q16, void* q8, int is_weight,
                                    int* sat, void* stream) {
    using namespace hupr;
    if (rows < 0 || ld <= 0 || ch < 0 || ch_off < 0 || ch_off + ch > ld || ld % 32 || ch % 32 || ch_off % 32) return HUPR_ERR_BAD_ARG;
    if (rows == 0 || ch == 0) return HUPR_OK;
    if (!hi || !q16 || !q8) return HUPR_ERR_BAD_ARG;
    if (((uintptr_t)hi | (uintptr_t)lo | (uintptr_t)q16 | (uintptr_t)q8) & 15) return HUPR_ERR_ALIGNMENT;
    if (int rc = device_check_sm100()) return rc;
    const int groups = ch / 8;
    const long long items = rows * groups;
    const int sms = device_sm_count();
    if (sms <= 0) return HUPR_ERR_CUDA;
    long long blocks = (items + 255) / 256;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    quantize_planes_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)hi, (const __nv_bfloat16*)lo, items, ld, ch_off,
                                                                                groups, (__half*)q16, (uint8_t*)q8, is_weight, sat);
    return pack_status();
}

extern "C" int hupr_broadcast_f32(const float* src, float* dst, int n, void* stream) {
    using namespace hupr;
    if (n < 0) return HUPR_ERR_BAD_ARG;
    if (n == 0) return HUPR_OK;
    if (!src || !dst) return HUPR_ERR_BAD_ARG;
    if (int rc = device_check_sm100()) return rc;
    broadcast_f32_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(src, dst, n);
    return pack_status();
}

extern "C" int hupr_gcn_bias_rows(const float* bias, int batch, int rows, void* rows_hi, void* rows_lo, void* stream) {
    using namespace hupr;
    if (batch < 0 || rows < batch * 14) return HUPR_ERR_BAD_ARG;
    if (rows == 0) return HUPR_OK;
    if (!bias || !rows_hi) return HUPR_ERR_BAD_ARG;
    if (int rc = device_check_sm100()) return rc;
    const long long threads = (long long)rows * 512;
    gcn_bias_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(bias, batch, rows, (__nv_bfloat16*)rows_hi,
                                                                                            (__nv_bfloat16*)rows_lo);
    return pack_status();
}

extern "C" int hupr_bump_i32(int* counter, void* stream) {
    using namespace hupr;
    if (!counter) return HUPR_ERR_BAD_ARG;
    if (int rc = device_check_sm100()) return rc;
    bump_i32_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter);
    return pack_status();
}

extern "C" int hupr_memset_zero(void* dst, size_t bytes, void* stream) {
    if (bytes == 0) return HUPR_OK;
    if (!dst) return HUPR_ERR_BAD_ARG;
    return cudaMemsetAsync(dst, 0, bytes, (cudaStream_t)stream) == cudaSuccess ? HUPR_OK : HUPR_ERR_CUDA;
}
