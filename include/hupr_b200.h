/* hupr_b200 — C ABI of the B200-native HuPR hot path (libhupr_b200.so).
 *
 * The reference (robert80203/HuPR) has no FFI layer: its hot path sits behind plain Python
 * signatures (SURVEY.md §8 b).  Each entry point below names the reference interface it replaces.
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer owned by the caller unless the name says `host`;
 *     the library never allocates, frees or retains caller memory;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, no internal
 *     synchronisation, CUDA-graph capturable;
 *   - return 0 on success, a negative HUPR_ERR_* code otherwise; nothing throws across the boundary;
 *   - there is no CPU fallback: on a non-sm_100 device the calls return HUPR_ERR_ARCH.
 */
#ifndef HUPR_B200_H
#define HUPR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HUPR_OK 0
#define HUPR_ERR_BAD_ARG (-1)
#define HUPR_ERR_ALIGNMENT (-2)
#define HUPR_ERR_CUDA (-3)
#define HUPR_ERR_ARCH (-4)
#define HUPR_ERR_WORKSPACE (-5)

/* Library version: major*10000 + minor*100 + patch. */
int hupr_version(void);

/* Human-readable text for an error code (static storage). */
const char* hupr_error_string(int code);

/* ---------------------------------------------------------------------------------------------
 * FFT cascade.  Replaces RadarObject.getadcDataFromDCA1000 + RadarObject.generateHeatmap
 *   /root/reference/preprocessing/process_iwr1843.py:54-83   (int16 2-lane IQ de-interleave)
 *   /root/reference/preprocessing/process_iwr1843.py:106-173 (TDM demux, clutter removal, range-Doppler
 *                                                             fft2, elevation+azimuth FFTs, crop/shift/flip)
 * adc  : int16 [n_frame_sensors][192 chirps][4 RX][256 samples] in DCA1000 2-lane order, i.e. per pair of
 *        samples the four words {I(2n), I(2n+1), Q(2n), Q(2n+1)}  (786 432 B per frame-sensor), 16-B aligned
 * cube : float2 (complex64) [n_frame_sensors][16 doppler][64 range][64 azimuth][8 elevation], 16-B aligned
 *        — the same axis order and index maps as the reference's complex128 .npy cube.
 */
int hupr_fft_cascade_i16(const int16_t* adc, void* cube, int n_frame_sensors, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Dense contraction on the tcgen05 tensor cores: implicit-GEMM convolution / GEMM with fused epilogue.
 * Replaces every nn.Conv3d / nn.Conv2d / einsum-bmm call of the model
 *   /root/reference/models/layers.py:24-32,45-63,81-95,116-123,126-133,195-210
 *
 *   out[pos, co] = act( scale[co] * sum_{tap,ci} A[pos + off(tap), a_ch_off + ci] * W[tap][co][ci] + shift[co] + R[pos, co] )
 *   act(v) = v > 0 ? v : slope[co] * v        (slope NULL = identity, 0 = ReLU, PReLU weight otherwise)
 *
 * Activations are channels-last bf16 "split" tensors: value = hi + lo (two bf16 planes of identical shape).  When the
 * lo planes are given, three tensor-core products (hi*hi + lo*hi + hi*lo) with fp32 accumulation give fp32-equivalent
 * results (the reference computes in fp32); with a_lo = w_lo = NULL a single bf16 product is used.
 * Filter taps: kd x kh x kw with paddings pd, ph, pw; H and W must be 'same' (k = 2p+1); D_out = D + 2pd - kd + 1.
 * Weights: bf16 [kd*kh*kw][cout][cin] (cin contiguous); with w_batched = 1 (taps must be 1) the first weight dim is the
 * sample index instead (per-sample B operand: the attention matmuls).
 * cin, cout must be multiples of 64 (pad with zero weights / channels); W must divide 128 or be a multiple of 128.
 */
typedef struct hupr_conv_desc {
    const void* a_hi; const void* a_lo;      /* [n][d][h][w][ca] bf16 */
    int n, d, h, w, ca;
    int a_ch_off, cin;                       /* channel slice of A that is contracted */
    const void* w_hi; const void* w_lo;      /* [taps | n][cout][cin] bf16 */
    int cout;
    int kd, kh, kw, pd, ph, pw;
    int w_batched;
    const float* scale; const float* shift; const float* slope;   /* [cout] each, may be NULL */
    const void* r_hi; const void* r_lo; int r_ld, r_ch_off;       /* residual [positions][r_ld] bf16 split, may be NULL */
    void* o_hi; void* o_lo; int o_ld, o_ch_off;                   /* bf16 split output [positions][o_ld], may be NULL */
    float* o_f32; int o_f32_ld;                                   /* fp32 output [positions][o_f32_ld], may be NULL */
} hupr_conv_desc;

int hupr_conv_gemm(const hupr_conv_desc* desc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HUPR_B200_H */
