// Fast layout changes of bf16 hi/lo tensors between channels-last ([position][channel]) and position-major ([channel][position]):
// 64 x 64 tiles, both planes in one launch, 16-byte global loads, 4-byte shared-memory reads, fully coalesced 128-byte row stores.
// Used by hupr_transpose_split (dense [n][s][c] -> [n][c][s]; V / K / Q / P operands of the attention GEMMs) and by hupr_to_kmajor /
// hupr_to_kmajor_multi (zero-padded position index of the weight-gradient GEMMs, several pre-shifted copies from ONE read).
#include "common.cuh"

namespace hupr {

struct TileTransposeParams {
    // source: element (pos, ch) at pos*ld + ch_off + ch
    int ld, ch_off;
    // destination addressing
    int dense;                 // 1: dst[(n*c_total + ch)*s + (pos - n*s)], s positions per sample
    int s, c_total;
    // padded position-major mode (dense == 0): P(pos) from the 4-D decomposition, copy k at dst + k*copy_stride, written at P - (shift0 + k)
    int d, h, w, dp, hp, wp, pd, ph, pw;
    int n_copies, shift0;
    long long copy_stride, ppad;
};

__global__ void __launch_bounds__(256)
tile_transpose_kernel(const uint16_t* __restrict__ src_hi, const uint16_t* __restrict__ src_lo, uint16_t* __restrict__ dst_hi,
                      uint16_t* __restrict__ dst_lo, const TileTransposeParams p) {
    pdl_wait();                  // programmatic dependent launch: the predecessor's writes are visible from here on (common.cuh)
    pdl_launch_dependents();     // the successor may start its prologue now; it waits the same way before touching memory
    __shared__ __align__(16) uint16_t tile[2][64][66];             // [plane][channel][position], pitch 66 keeps 4-byte reads conflict free
    const long long pos0 = (long long)blockIdx.x * 64;
    const int c0 = blockIdx.y * 64;
    const int tid = threadIdx.x;
    const int planes = src_lo ? 2 : 1;
    {   // load: thread = (position tid/4, 16-channel quarter tid%4)
        const int pr = tid >> 2, q = tid & 3;
        const size_t off = (size_t)(pos0 + pr) * p.ld + p.ch_off + c0 + q * 16;
        for (int pl = 0; pl < planes; ++pl) {
            const uint4* g = reinterpret_cast<const uint4*>((pl ? src_lo : src_hi) + off);
            const uint4 a = __ldg(g), b = __ldg(g + 1);
            const uint32_t wds[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                tile[pl][q * 16 + 2 * i][pr] = (uint16_t)(wds[i] & 0xFFFFu);
                tile[pl][q * 16 + 2 * i + 1][pr] = (uint16_t)(wds[i] >> 16);
            }
        }
    }
    __syncthreads();
    // store: warp w owns channels 8w .. 8w+7; lane owns positions 2*lane, 2*lane + 1 (adjacent in the destination)
    const int warp = tid >> 5, lane = tid & 31;
    const long long pos = pos0 + 2 * lane;
    long long base;                                   // destination index of (channel 0, this position), copy 0, before the shift
    if (p.dense) {
        const long long n = pos / p.s;
        base = n * (long long)p.c_total * p.s + (pos - n * p.s);
    } else {
        long long t = pos;
        const int w = (int)(t % p.w); t /= p.w;
        const int h = (int)(t % p.h); t /= p.h;
        const int d = (int)(t % p.d);
        const long long n = t / p.d;
        base = ((n * p.dp + d + p.pd) * p.hp + h + p.ph) * (long long)p.wp + w + p.pw;
    }
    const long long row_stride = p.dense ? p.s : p.ppad;
    for (int pl = 0; pl < planes; ++pl) {
        uint16_t* dst = pl ? dst_lo : dst_hi;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int ch = warp * 8 + i;
            const uint32_t v = *reinterpret_cast<const uint32_t*>(&tile[pl][ch][2 * lane]);
            for (int k = 0; k < p.n_copies; ++k) {
                const long long o = k * p.copy_stride + (long long)(c0 + ch) * row_stride + base - (p.shift0 + k);
                if ((o & 1) == 0) {
                    *reinterpret_cast<uint32_t*>(dst + o) = v;
                } else {
                    dst[o] = (uint16_t)(v & 0xFFFFu);
                    dst[o + 1] = (uint16_t)(v >> 16);
                }
            }
        }
    }
}

static int tt_check_sm100() {
    return device_check_sm100();      // cached per device (capi.cu)
}

// Dense [n][s][ld] (channels ch_off .. +c) -> [n][c][s].  Caller guarantees s % 64 == 0, c % 64 == 0, 16-byte aligned rows.
int fast_transpose_dense(const void* in_hi, const void* in_lo, int n, int s, int c, int in_ld, int in_ch_off, void* out_hi, void* out_lo,
                         cudaStream_t stream) {
    int rc = tt_check_sm100();
    if (rc != HUPR_OK) return rc;
    TileTransposeParams p = {};
    p.ld = in_ld; p.ch_off = in_ch_off; p.dense = 1; p.s = s; p.c_total = c; p.n_copies = 1; p.shift0 = 0; p.copy_stride = 0; p.ppad = 0;
    const dim3 grid((unsigned)((long long)n * s / 64), c / 64);
    launch_k(tile_transpose_kernel, dim3(grid), dim3(256), (size_t)(0), stream, (const uint16_t*)in_hi, (const uint16_t*)in_lo, (uint16_t*)out_hi, (uint16_t*)out_lo, p);
    note_launches(1);
    return cudaGetLastError() == cudaSuccess ? HUPR_OK : HUPR_ERR_CUDA;
}

int fast_to_kmajor(const void* src_hi, const void* src_lo, int n, int d, int h, int w, int ld, int ch_off, int c, void* dst_hi, void* dst_lo,
                   int dp, int hp, int wp, int pd, int ph, int pw, int shift0, int n_copies, long long copy_stride, long long ppad,
                   cudaStream_t stream) {
    int rc = tt_check_sm100();
    if (rc != HUPR_OK) return rc;
    TileTransposeParams p = {};
    p.ld = ld; p.ch_off = ch_off; p.dense = 0; p.s = 0; p.c_total = c;
    p.d = d; p.h = h; p.w = w; p.dp = dp; p.hp = hp; p.wp = wp; p.pd = pd; p.ph = ph; p.pw = pw;
    p.n_copies = n_copies; p.shift0 = shift0; p.copy_stride = copy_stride; p.ppad = ppad;
    const dim3 grid((unsigned)((long long)n * d * h * w / 64), c / 64);
    launch_k(tile_transpose_kernel, dim3(grid), dim3(256), (size_t)(0), stream, (const uint16_t*)src_hi, (const uint16_t*)src_lo, (uint16_t*)dst_hi, (uint16_t*)dst_lo, p);
    note_launches(1);
    return cudaGetLastError() == cudaSuccess ? HUPR_OK : HUPR_ERR_CUDA;
}

}  // namespace hupr

extern "C" int hupr_to_kmajor_multi(const void* src_hi, const void* src_lo, int n, int d, int h, int w, int ld, int ch_off, int c, void* dst_hi,
                                    void* dst_lo, int dp, int hp, int wp, int pd, int ph, int pw, int first_shift, int n_copies,
                                    long long rows_per_copy, long long ppad, void* stream) {
    using namespace hupr;
    if (!src_hi || !dst_hi || n <= 0 || d <= 0 || h <= 0 || w <= 0 || c <= 0 || c % 64 || ch_off < 0 || ch_off % 8 || ch_off + c > ld || ld % 8)
        return HUPR_ERR_BAD_ARG;
    if ((src_lo == nullptr) != (dst_lo == nullptr) || n_copies < 1 || n_copies > 8 || rows_per_copy < c) return HUPR_ERR_BAD_ARG;
    const long long positions = (long long)n * d * h * w;
    if (positions % 64 || w % 2 || d + pd > dp || h + ph > hp || w + pw > wp || pd < 0 || ph < 0 || pw < 0) return HUPR_ERR_BAD_ARG;
    if (ppad < (long long)n * dp * hp * wp || positions / 64 > 2147483647LL) return HUPR_ERR_BAD_ARG;
    if ((long long)(pd * hp + ph) * wp + pw - (first_shift + n_copies - 1) < 0 || first_shift < -pw - 1) return HUPR_ERR_BAD_ARG;
    if (((uintptr_t)src_hi | (uintptr_t)src_lo) & 15 || ((uintptr_t)dst_hi | (uintptr_t)dst_lo) & 3 || ppad % 2) return HUPR_ERR_ALIGNMENT;
    return fast_to_kmajor(src_hi, src_lo, n, d, h, w, ld, ch_off, c, dst_hi, dst_lo, dp, hp, wp, pd, ph, pw, first_shift, n_copies,
                          rows_per_copy * ppad, ppad, (cudaStream_t)stream);
}
