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

/* Number of kernels this library has launched in the calling process (all entry points, all streams). */
long long hupr_launch_count(void);

/* Programmatic dependent launch for the latency-bound inference chain (convolutions, attention, bridge / GCN / argmax kernels): when on
 * (default off; environment HUPR_PDL=1 turns it on) those kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization and
 * wait with griddepcontrol.wait before their first global access, so a kernel's launch and prologue overlap its predecessor's tail (also
 * as programmatic edges inside captured CUDA graphs).  Returns the previous setting; a negative argument only queries. */
int hupr_set_pdl(int on);

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
 * sample index instead (per-sample B operand: the attention matmuls).  With w_ld != 0 the weight rows are w_ld elements apart and
 * start at column w_ch_off (a channel slice of a wider channels-last tensor used as the B operand).
 * cin, cout must be multiples of 64 (pad with zero weights); the contracted slice may run up to 63 channels past `ca` —
 * those channels read as zero (TMA out-of-bounds fill), e.g. ca = 32, cin = 64.  W must divide 128 or be a multiple of 128.
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
    int w_ld, w_ch_off;                                           /* weight row stride / first column in elements (0, 0 = dense [..][cout][cin]) */
    long long a_n_stride;                                         /* elements between samples of A (0 = dense d*h*w*ca); smaller values give
                                                                     overlapping sliding windows over a frame stream */
    int k_split;                                                  /* > 1: split the contraction over this many CTAs per output tile and ADD the
                                                                     partial sums atomically into o_f32 (pre-zeroed by the caller; no scale / shift /
                                                                     slope / residual / bf16 output).  For contractions with few output tiles and a very
                                                                     long K: the weight-gradient GEMMs (K = positions) */
    int w_k_off;                                                  /* added to the weight operand's contracted-axis coordinate (may be negative; out of
                                                                     range reads zero): shifted correlations over position-major operands */
    const float* row_vec; int row_mode;                           /* per-POSITION vector [positions] applied to the accumulator before the rest of the
                                                                     epilogue (attention backward, layers.py:126-133 under autograd; no scale/shift/slope):
                                                                     1: out = exp(acc - row_vec[pos])          P rebuilt from the logits and the saved log-sum-exp
                                                                     2: out = r * (acc - row_vec[pos])         dS = P * (dP - rowdot): r (the "residual" operand) MULTIPLIES */
    void* ws; size_t ws_bytes;                                    /* optional scratch (16-B aligned) enabling COOPERATIVE split-K when the output has too few
                                                                     tiles to fill the GPU (batch-1 inference, the PRGCN GEMMs): slices of the contraction run on
                                                                     separate CTAs, park fp32 partial tiles here, and the last arrival adds them in slice order
                                                                     (deterministic) and runs the full epilogue — one launch, any epilogue.  The first 4096 bytes
                                                                     are arrival counters: zero them once; every launch leaves them zero.  One workspace must not
                                                                     be shared by launches that may run concurrently (different streams) */
    int no_tma_store;                                             /* 1: keep the thread-per-row global stores even where the shared-memory-staged TMA-store
                                                                     epilogue applies (row tiles: w a multiple of 128, bf16 split output) — for A/B measurements */
    int nprod;                                                    /* tensor-core products per k-step: 0 = by operands (3 when the lo planes are given, else 1);
                                                                     1 = hi*hi only although lo planes exist (they are not read): plain bf16 compute with fp32
                                                                     accumulation, the precision BASELINE.json configs[3] names for training; 3 = require lo planes;
                                                                     2 = the quantised-cross-term arithmetic below where the shape allows it (else 3) */
    double* stats; int stats_ld;                                  /* optional: per-output-channel sums of the epilogue values over all positions, ADDED to
                                                                     stats[co] (sum v) and stats[stats_ld + co] (sum v^2); zero-filled by the caller.  Fuses the
                                                                     batch statistics of a train-mode nn.BatchNorm3d (/root/reference/models/layers.py:45-53)
                                                                     into the convolution that produces its input */
    /* Two-unit arithmetic for the 3-tap convolutions (nprod == 2): with x = x16 + xl, the product a*w is evaluated as
     *   a16*w16 (fp16 operands, full tensor rate)  +  al*w  +  a*wl  (both cross terms as e4m3 x e4m3 products at twice the rate)
     * into ONE fp32 accumulator: 1 + 1/2 + 1/2 = 2 tensor units per k-step instead of the 3 bf16 products, ~2^-13 relative per product
     * (measured whole-network error: DESIGN.md §2).  The operands carry fixed power-of-two scales so that every product is scaled by
     * 2^16 (undone exactly in the epilogue):  a16 = fp16(a * 2^2), a8 = e4m3(a * 2^1), al8 = e4m3((a - a16) * 2^12);
     * w16 = fp16(w * 2^14), w8 = e4m3(w * 2^4), wl8 = e4m3((w - w16) * 2^15).  Conversions saturate: beyond |a| = 224 the cross terms
     * lose accuracy gradually (a term then has single-fp16-product accuracy, 2^-12); the hard limits are |a| < 16 376 and |w| < 3.99 (fp16
     * planes).  Two planes per tensor, with the geometry of a_hi / w_hi (same ca, a_ch_off, w_ld, ... all multiples of 32 here):
     *   *_q16 : fp16 [rows][ld]
     *   *_q8  : e4m3 bytes [rows][ld / 32][2][32] — per 32-channel block the 32 values x8 followed by the 32 residuals xl8, i.e. 64-byte
     *           rows that the tensor core reads as two K = 32 slices (32-byte rows under SWIZZLE_32B read at half the rate)
     * hupr_quantize_planes writes them, or a producing hupr_conv_gemm call does through o_q*.  All NULL = not available. */
    const void* a_q16; const void* a_q8;
    const void* w_q16; const void* w_q8;
    void* o_q16; void* o_q8;                                      /* optional: ALSO store the output as activation planes of that form ([positions][o_ld]
                                                                     at o_ch_off like o_hi), for a following nprod == 2 convolution; needs o_hi */
    int* q_sat;                                                   /* optional device counter, incremented whenever a value written to o_q16 lies outside
                                                                     the fp16 plane's range (|a| >= 16 376): the planes then no longer represent the tensor */
} hupr_conv_desc;

int hupr_conv_gemm(const hupr_conv_desc* desc, void* stream);

/* 1 when hupr_conv_gemm would run `desc` (with nprod == 2 and its quantised operand planes) on the two-unit kernel, 0 when it would fall
 * back to the three bf16 products (shape not handled there), negative = HUPR_ERR_*.  Lets a caller skip producing planes nobody reads. */
int hupr_conv_quant_eligible(const hupr_conv_desc* desc);

/* The operand planes of the two-unit arithmetic from a bf16 split tensor: rows x [ch_off, ch_off + ch) of [rows][ld] (ld, ch_off, ch
 * multiples of 32).  is_weight selects the scale set (0: activation 2^2 / 2^1 / 2^12, 1: weight 2^14 / 2^4 / 2^15).  q16: fp16 [rows][ld];
 * q8: e4m3 bytes [rows][ld / 32][2][32] (hupr_conv_desc).  lo may be NULL.  sat: optional device counter incremented when a value
 * lies outside the fp16 plane's range (hupr_conv_desc.q_sat). */
int hupr_quantize_planes(const void* hi, const void* lo, long long rows, int ld, int ch_off, int ch, void* q16, void* q8, int is_weight,
                         int* sat, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused spatial attention (flash-style: the [S, S] logits never leave the SM).  Replaces
 * MultiScaleCrossSelfAttentionPRGCN.attention  /root/reference/models/layers.py:126-133 (+ the residual add of :146,148)
 *   out[b, n, :] = sum_m softmax_m( <Q[b, n, :], K[b, m, :]> ) * V[b, m, :]  (+ R[b, n, :])        no 1/sqrt(d) scale
 * q, k : bf16 split, rows [batch][s][q_ld | k_ld], the c contracted channels start at q_off | k_off
 * vt   : bf16 split [batch][c][s]  (V transposed: hupr_transpose_split)
 * r    : optional residual rows [batch][s][r_ld] at r_off;  o: output rows [batch][s][o_ld] at o_off (bf16 split)
 * Supported: c == 64 or 128, s a multiple of 128, all four lo planes present (other shapes: HUPR_ERR_BAD_ARG — compose
 * hupr_conv_gemm(w_batched) + hupr_softmax_rows instead).
 */
typedef struct hupr_attn_desc {
    const void* q_hi; const void* q_lo; int q_ld, q_off;
    const void* k_hi; const void* k_lo; int k_ld, k_off;
    const void* vt_hi; const void* vt_lo;
    const void* r_hi; const void* r_lo; int r_ld, r_off;
    void* o_hi; void* o_lo; int o_ld, o_off;
    int batch, s, c;
    float* lse;      /* optional float [batch][s]: log sum_m exp(<Q[n], K[m]>) per query row — lets the backward pass rebuild
                        P = exp(logits - lse) in a GEMM epilogue (hupr_conv_desc.row_mode) instead of a separate softmax; NULL = not written */
} hupr_attn_desc;

int hupr_attention_fwd(const hupr_attn_desc* desc, void* stream);

/* Fused backward of the attention above for c == 64 (no [S, S] matrix in memory): given the forward's operands, the gradient dO of its
 * output, the saved log-sum-exp rows (hupr_attn_desc.lse) and rowdot[b][n] = <dO[n], (P V)[n]> (hupr_rowdot on the forward output minus
 * the residual), ADDS  dQ = dS K,  dK = dS^T Q,  dV = P^T dO  to fp32 buffers, where P = exp(Q K^T - lse), dS = P o (dO V^T - rowdot)
 * (autograd of /root/reference/models/layers.py:126-133).  q, k, v, do: bf16 split rows [batch][s][*_ld] with the 64 channels at *_off
 * (v is NOT transposed here); dq, dk, dv: float [batch][s][*_ld] at *_off, zero-filled (or holding earlier contributions) by the caller.
 * s a multiple of 128.  Called by hupr_b200.training.TrainStep for the head-dim-64 level. */
typedef struct hupr_attn_bwd_desc {
    const void* q_hi; const void* q_lo; int q_ld, q_off;
    const void* k_hi; const void* k_lo; int k_ld, k_off;
    const void* v_hi; const void* v_lo; int v_ld, v_off;
    const void* do_hi; const void* do_lo; int do_ld, do_off;
    const float* lse; const float* rowdot;
    float* dq; int dq_ld, dq_off;
    float* dk; int dk_ld, dk_off;
    float* dv; int dv_ld, dv_off;
    int batch, s, c;
} hupr_attn_bwd_desc;

int hupr_attention_bwd(const hupr_attn_bwd_desc* desc, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Loader bridge.  Replaces Normalize.__call__ + the window assembly of HuPR3D_horivert.__getitem__
 *   /root/reference/datasets/base.py:13-24, /root/reference/datasets/dataset.py:120-150
 * cube    : float2 [n_frame_sensors][16][64][64][8] (output of hupr_fft_cascade_i16)
 * slot_fs : int32 [n_slots]  frame-sensor index feeding window slot s = sample*8 + group_frame (the clamped window
 *           of dataset.py:126-138 is computed by the caller; it is an integer index map)
 * vrdae   : float [n_slots][8 kept Doppler rows 4..11][2 re,im][64 range][64 azimuth][8 elevation]; every
 *           (slot, row, re|im, elevation) plane is standardised over its 64x64 cells: (x - mean) / unbiased std.
 */
int hupr_window_normalize(const void* cube, const int32_t* slot_fs, int n_slots, float* vrdae, void* stream);

/* Chirp encoder.  Replaces HuPRNet.forward_chirp (one sensor) + MNet.forward
 *   /root/reference/models/networks.py:23-33, /root/reference/models/chirp_networks.py:11-21
 * vrdae  : float [n_slots][8][2][64][64][8];  weight: float [32][2][2] (temporalConvWx1x1.weight), bias: float [32]
 * out    : bf16 split channels-last [n_slots][64][64][32]  (= [B][D=8][H][W][32]); out_lo may be NULL.
 */
int hupr_mnet_fwd(const float* vrdae, const float* weight, const float* bias, void* out_hi, void* out_lo, int n_slots, void* stream);

/* Streaming variant of the two entry points above: standardisation and MNet are per-frame operations and
 * consecutive windows share 7 of 8 frames, so a stream computes them once per frame-sensor.  Same arithmetic as
 * hupr_window_normalize + hupr_mnet_fwd (datasets/base.py:13-24, models/networks.py:23-33, models/chirp_networks.py:11-21).
 *   hupr_plane_stats     cube float2 [n][16][64][64][8] -> workspace: mean and 1/std of the 16 (re|im, elevation) planes of the 8
 *                        kept Doppler rows of every frame-sensor (hupr_frame_features_workspace_bytes(n) bytes)
 *   hupr_frame_features  frame-sensors [first, first+n) -> bf16 split features [n][64][64][32] with one sensor's MNet weights
 * A window [B][8][64][64][32] is then an overlapping strided view of the feature buffer (hupr_conv_desc.a_n_stride). */
size_t hupr_frame_features_workspace_bytes(int n_frame_sensors);
int hupr_plane_stats(const void* cube, int n_frame_sensors, void* workspace, size_t ws_bytes, void* stream);
int hupr_frame_features(const void* cube, const void* stats, int first_frame_sensor, int n_frame_sensors, const float* weight,
                        const float* bias, void* out_hi, void* out_lo, void* stream);

/* Bi/tri-linear resampling with align_corners=True on split channels-last tensors.  Replaces nn.Upsample / F.interpolate
 *   /root/reference/models/layers.py:84,89,199,204
 * in  : [n][di][hi][wi][in_ld] channels in_ch_off .. +c ;  out : [n][dout][ho][wo][out_ld] channels out_ch_off .. +c
 */
int hupr_resample_linear(const void* in_hi, const void* in_lo, int n, int di, int hi, int wi, int c, int in_ld, int in_ch_off,
                         void* out_hi, void* out_lo, int dout, int ho, int wo, int out_ld, int out_ch_off, void* stream);

/* Row softmax of the attention logits (/root/reference/models/layers.py:131; rows = queries, columns = keys).
 * logits: float [rows][cols] -> p: bf16 split [rows][cols]; cols multiple of 4, <= 4096. */
int hupr_softmax_rows(const float* logits, void* p_hi, void* p_lo, long long rows, int cols, void* stream);

/* [n][s][in_ld](channels in_ch_off..+c) -> [n][c][s]: V operand of the second attention matmul (layers.py:132). */
int hupr_transpose_split(const void* in_hi, const void* in_lo, int n, int s, int c, int in_ld, int in_ch_off,
                         void* out_hi, void* out_lo, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Heatmap head + pose-refinement GCN.  Replaces PRGCN.forward and the final sigmoid of HuPRNet.forward
 *   /root/reference/models/gcn_networks.py:6-64, /root/reference/models/layers.py:97-112, /root/reference/models/networks.py:40
 * logits      : float channels-last [batch][64*64][ld], first 14 channels are the keypoint logits
 * weights[l]  : float [1024][1024] (gcn.L{1,2,3}.weight), biases[l]: float [1024][14]; adj: float [14][14]
 * heatmap     : float [batch][14][64][64] = sigmoid(logits);  gcn_heatmap: float [batch][14][64][64]
 * workspace   : hupr_prgcn_workspace_bytes(batch) bytes, 16-B aligned.  weights/biases are HOST arrays of 3 device pointers.
 */
size_t hupr_prgcn_workspace_bytes(int batch);
int hupr_prgcn_fwd(const float* logits, int ld, const float* const* weights, const float* const* biases, const float* adj,
                   void* workspace, size_t ws_bytes, float* heatmap, float* gcn_heatmap, int batch, void* stream);

/* Tensor-core formulation of the same PRGCN (the caller runs the three 1024x1024 contractions with hupr_conv_gemm in the transposed
 * layout Yt[(b,j), q] = sum_p St[(b,j), p] W[q, p], bias as residual rows, ReLU as slope):
 *   hupr_gcn_nodes  logits -> heatmap = sigmoid(logits) [batch][14][64][64]  and  St0 = (X.A)^T bf16 split rows [(b,j)][1024]
 *   hupr_gcn_mix    St_next[(b,j'), q] = sum_j A[j,j'] Yt[(b,j), q]   (bf16 split in and out, rows [(b,j)][1024])
 *   hupr_gcn_heads  Yt_3 float rows [(b,j)][y_ld] (= [batch][14][32][32] maps) -> bilinear x2 + sigmoid -> [batch][14][64][64] */
int hupr_gcn_nodes(const float* logits, int ld, const float* adj, float* heatmap, void* s_hi, void* s_lo, int batch, void* stream);
int hupr_gcn_mix(const void* y_hi, const void* y_lo, const float* adj, void* s_hi, void* s_lo, int batch, void* stream);
int hupr_gcn_heads(const float* y, int y_ld, float* gcn_heatmap, int batch, void* stream);

/* Keypoint decode.  Replaces get_max_preds (/root/reference/misc/metrics.py:10-38).
 * maps: float [n_maps][64*64] -> preds float [n_maps][2] = (x, y) of the first maximum, zeroed where max <= 0; maxvals may be NULL. */
int hupr_keypoints_argmax(const float* maps, int n_maps, float* preds, float* maxvals, void* stream);

/* Evaluation tail on the device.  Object keypoint similarity of one predicted pose against the one ground-truth
 * pose of its image — COCOeval.computeOks of the reference's vendored evaluator (/root/reference/misc/cocoeval.py:192-236) with all k
 * joints labelled visible, as /root/reference/datasets/base.py:60 writes them:
 *   oks[i] = mean_j exp(-((dx^2 + dy^2) / (2 sigma_j)^2 / (area[i] + eps) / 2)),  eps = np.spacing(1); float64 like the original.
 * pred, gt: float [n][k][2] image-pixel (x, y); area, sigmas: double; per_joint (optional): double [n][k] = the k exponentials
 * (per-joint evaluation, cocoeval.py:232-233). */
int hupr_keypoint_oks(const float* pred, const float* gt, const double* area, const double* sigmas, int n, int k, double* oks,
                      double* per_joint, void* stream);

/* Heatmap loss.  Replaces LossComputer.computeLoss + generateTarget
 *   /root/reference/misc/losses.py:23-48, /root/reference/misc/utils.py:6-65
 * heatmap, gcn_heatmap: float [batch][14][64][64] (post-sigmoid); joints: int64 [batch][14][2] image pixels (256-px frame)
 * losses : float [3] = {loss1 + loss2, loss2, loss1} (mean BCE, logs clamped at -100)
 * targets: optional float [batch][14][64][64] Gaussian targets; gt2d: optional float [batch][14][2] target argmax (x, y)
 * workspace: batch*14*2 doubles. */
int hupr_heatmap_loss_fwd(const float* heatmap, const float* gcn_heatmap, const long long* joints, int batch,
                          void* workspace, size_t ws_bytes, float* losses, float* targets, float* gt2d, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training building blocks (hupr_b200/training.py assembles them into the whole backward pass).
 *
 * hupr_heatmap_loss_bwd: gradient of loss = w1*BCE(heatmap, T) + w2*BCE(gcn_heatmap, T) (mean reduction) with respect to the
 * pre-sigmoid values — what autograd computes for /root/reference/misc/losses.py:23-48 + the sigmoids of models/networks.py:40 and
 * models/gcn_networks.py:64.  d_heat_logits: float channels-last [batch][4096][ld] (first 14 channels written; the layout of the head
 * convolution's output), d_gcn_pre: float [batch][14][64][64].
 *
 * hupr_adam_step: one torch.optim.Adam step with coupled L2 weight decay (tools/base.py:47: lr 1e-4, betas (0.9, 0.999), eps 1e-8,
 * weight_decay 1e-4) over a flat fp32 buffer of n elements; `step` is the 1-based step count used for the bias corrections. */
/* Channels-last bf16 split view used by the training kernels: element (pos, ch) lives at index pos*ld + ch_off + ch of both planes
 * (lo may be NULL for single-bf16 tensors). */
typedef struct hupr_tensor_view { void* hi; void* lo; int ld, ch_off; } hupr_tensor_view;

/* Training-mode element-wise / reduction kernels (what autograd + nn.BatchNorm3d(training) + nn.PReLU + F.interpolate backward do around
 * the convolutions of /root/reference/models/layers.py:8-70,186-217 and the softmax of :131).  `positions` rows of `c` channels each.
 *   hupr_channel_sums  s1[ch] += sum f1, s2[ch] += sum f2 (double, caller zero-fills):
 *        mode 0 STATS   a = z                      f1 = z,  f2 = z^2                      BatchNorm batch statistics
 *        mode 1 BN_BWD  a = g, b = z, mask = y     f1 = g', f2 = g' * (z - mean) * rstd   g' = g * [y > 0] (mask optional)
 *        mode 2 PRELU   a = g, b = s (optional)    f1 = g * s * [s <= 0], f2 = g          PReLU slope gradient / bias gradient
 *   hupr_affine_act    out = act(scale1*z + shift1 (+ scale2*r + shift2)); act(v) = v > 0 ? v : slope[ch]*v  (NULL scale = 1, shift = 0,
 *                      slope NULL = identity)
 *   hupr_bn_bwd_apply  out = k1 * (g' - k2 - (z - mean)*rstd * k3)          k1 = gamma*rstd, k2 = mean(g'), k3 = mean(g' zhat)
 *   hupr_act_bwd       out = g * (s > 0 ? 1 : slope[ch])
 *   hupr_accumulate    out = a + b + f   (a, b bf16 split views, f float [positions][f_ld] at f_off; each optional)
 *   hupr_resample_linear_bwd  adjoint of hupr_resample_linear: din float [n][di][hi][wi][in_ld] (+= , caller zero-fills)
 *   hupr_softmax_bwd_rows     ds = p * (dp - sum_m p*dp) per row (p bf16 split, dp float, ds bf16 split; cols % 4 == 0, <= 4096) */
int hupr_channel_sums(int mode, const hupr_tensor_view* a, const hupr_tensor_view* b, const hupr_tensor_view* mask, const float* mean,
                      const float* rstd, long long positions, int c, double* s1, double* s2, void* stream);
int hupr_affine_act(const hupr_tensor_view* z, const float* scale1, const float* shift1, const hupr_tensor_view* r, const float* scale2,
                    const float* shift2, const float* slope, const hupr_tensor_view* out, long long positions, int c, void* stream);
int hupr_bn_bwd_apply(const hupr_tensor_view* g, const hupr_tensor_view* z, const hupr_tensor_view* mask, const float* mean, const float* rstd,
                      const float* k1, const float* k2, const float* k3, const hupr_tensor_view* out, long long positions, int c, void* stream);
int hupr_act_bwd(const hupr_tensor_view* g, const hupr_tensor_view* s, const float* slope, const hupr_tensor_view* out, long long positions,
                 int c, void* stream);
/* Per-channel BatchNorm bookkeeping (nn.BatchNorm3d in train mode, /root/reference/models/layers.py:45-53, eps 1e-5, momentum 0.1) in one
 * launch each:
 *   hupr_bn_finalize      double sums s1 = sum z, s2 = sum z^2 over `count` positions -> mean, rstd = 1/sqrt(biased var + eps), the fused
 *                         affine scale = gamma*rstd, shift = beta - mean*scale, running_mean/var (unbiased variance) += momentum * (new - old),
 *                         num_batches_tracked += 1 (the three running-state pointers may be NULL)
 *   hupr_bn_bwd_finalize  double sums t1 = sum g', t2 = sum g' zhat -> k2 = t1/count, k3 = t2/count (hupr_bn_bwd_apply), dgamma = t2, dbeta = t1 */
int hupr_bn_finalize(const double* s1, const double* s2, long long count, const float* gamma, const float* beta, float eps, float momentum,
                     float* mean, float* rstd, float* scale, float* shift, float* running_mean, float* running_var,
                     long long* num_batches_tracked, int c, void* stream);
int hupr_bn_bwd_finalize(const double* t1, const double* t2, long long count, float* k2, float* k3, float* dgamma, float* dbeta, int c,
                         void* stream);
/* out[pos] = sum_ch a[pos][ch] * (b[pos][ch] - c[pos][ch])   (c optional) over `c_n` channels: the softmax-backward row term
 * rowsum(P o dP) = <dO, O - residual> of the attention backward. */
int hupr_rowdot(const hupr_tensor_view* a, const hupr_tensor_view* b, const hupr_tensor_view* c, float* out, long long positions, int c_n,
                void* stream);
int hupr_accumulate(const hupr_tensor_view* a, const hupr_tensor_view* b, const float* f, int f_ld, int f_off, const hupr_tensor_view* out,
                    long long positions, int c, void* stream);
int hupr_resample_linear_bwd(const hupr_tensor_view* g, int n, int dout, int ho, int wo, int c, float* din, int di, int hi, int wi, int in_ld,
                             int in_ch_off, void* stream);
int hupr_softmax_bwd_rows(const void* p_hi, const void* p_lo, const float* dp, void* ds_hi, void* ds_lo, long long rows, int cols, void* stream);

/* Backward pieces of the PRGCN head and of MNet (autograd of /root/reference/models/gcn_networks.py:23-64, chirp_networks.py:17-21):
 *   hupr_gcn_bias_grad  dbias[q][j] = sum_b dYt[(b,j)][q]                      (dYt bf16 split rows [(b,j)][1024]; dbias float [1024][14])
 *   hupr_gcn_heads_bwd  adjoint of hupr_gcn_heads' bilinear x2: d_pre float [batch][14][64][64] -> dyt float rows [(b,j)][1024] (+=)
 *   hupr_gcn_nodes_bwd  adjoint of hupr_gcn_nodes' node path: dSt0 rows -> d_logits float channels-last [batch][4096][ld] (+=)
 *   hupr_mnet_bwd       weight [32][2][2] / bias [32] gradients of MNet (double, +=) from the chirp-feature gradient [n_slots][4096][32] */
int hupr_gcn_bias_grad(const void* g_hi, const void* g_lo, float* dbias, int batch, void* stream);
int hupr_gcn_heads_bwd(const float* d_pre, float* dyt, int batch, void* stream);
int hupr_gcn_nodes_bwd(const void* s_hi, const void* s_lo, const float* adj, float* d_logits, int ld, int batch, void* stream);
int hupr_mnet_bwd(const float* vrdae, const float* weight, const float* bias, const void* g_hi, const void* g_lo, double* dweight, double* dbias,
                  int n_slots, void* stream);

/* Position-major copy for the weight-gradient GEMMs: src bf16 split [n][d][h][w][ld] channels ch_off..+c  ->  dst [c][ppad] with
 * P(n,d,h,w) = ((n*dp + d + pd)*hp + h + ph)*wp + w + pw.  The caller zero-fills dst once (padding cells are never written); a filter
 * tap is then the constant shift ((kd-pd)*hp + (kh-ph))*wp + (kw-pw) passed as hupr_conv_desc.w_k_off, and
 *   dW[tap][co][ci] = sum_P dYt[co][P] * Xt[ci][P + off]   is   hupr_conv_gemm(A = dYt, W = Xt, k_split, w_k_off)
 * (what autograd computes for the weights of nn.Conv3d / nn.Conv2d). */
int hupr_to_kmajor(const void* src_hi, const void* src_lo, int n, int d, int h, int w, int ld, int ch_off, int c,
                   void* dst_hi, void* dst_lo, int dp, int hp, int wp, int pd, int ph, int pw, int shift, long long ppad, void* stream);
/* Same re-layout, n_copies copies from ONE read of src: copy k is written at rows [k*rows_per_copy, +c) of dst with shift
 * first_shift + k (the kw = 0..2 operands of a 3-wide filter: first_shift = -pw, n_copies = 3).  Needs c % 64 == 0 and an even w. */
int hupr_to_kmajor_multi(const void* src_hi, const void* src_lo, int n, int d, int h, int w, int ld, int ch_off, int c, void* dst_hi,
                         void* dst_lo, int dp, int hp, int wp, int pd, int ph, int pw, int first_shift, int n_copies,
                         long long rows_per_copy, long long ppad, void* stream);
/* Weight gradient of a stride-1 convolution read straight from the channels-last tensors (no position-major copies):
 *   dW[tap][ci][co] += sum_{n,d,h,w} X[n, d+kd-pd, h+kh-ph, w+kw-pw, x_ch_off+ci] * dY[n, d, h, w, y_ch_off+co]      tap = (kd*KH + kh)*KW + kw
 * — what autograd derives for nn.Conv3d / nn.Conv2d weights in loss.backward() (/root/reference/tools/run.py:78; the convolutions of
 * /root/reference/models/layers.py:24-32,45-63,116-123,195-210).  The contraction runs over POSITIONS on the tensor cores with both
 * operands MN-major (channels contiguous), a filter tap being a TMA coordinate offset with out-of-bounds zero fill.
 * x  : bf16 split [n][d][h][w][cx];  dy : bf16 split [n][d_out][h][w][cy], d_out = d + 2*pd - kd + 1 (H, W are 'same': k = 2p+1)
 * dw : float [taps*cin][dw_ld], row = tap*cin + ci, column = co; ACCUMULATED with atomic adds (split-K) — zero it first.
 * cin, cout multiples of 64 (channel slices may run past cx / cy up to the next multiple of 64: those read as zero);
 * w must divide 64 or be a multiple of 64, h a multiple of 64 / min(w, 64).
 * batched != 0: the n samples are independent problems (no sum over n), dw[sample] = dw + sample * dw_batch_stride floats — the
 * "A^T B" products of the attention backward (/root/reference/models/layers.py:126-133 under autograd): dK = dS^T Q and dV = P^T dO are
 * contractions over the QUERY axis of row-major [queries][keys] matrices, i.e. x = the matrix with cx = keys "channels". */
typedef struct hupr_wgrad_desc {
    const void* x_hi; const void* x_lo; int n, d, h, w, cx, x_ch_off, cin;
    const void* dy_hi; const void* dy_lo; int cy, y_ch_off, cout;
    int kd, kh, kw, pd, ph, pw;
    float* dw; int dw_ld;
    int batched; long long dw_batch_stride;
    int nprod;                               /* 0 / 3: three products per k-step (fp32-equivalent); 1: hi*hi only (the lo planes are not read) */
} hupr_wgrad_desc;
int hupr_conv_wgrad(const hupr_wgrad_desc* desc, void* stream);

int hupr_heatmap_loss_bwd(const float* heatmap, const float* gcn_heatmap, const long long* joints, int batch, int ld, float w1, float w2,
                          float* d_heat_logits, float* d_gcn_pre, void* stream);
/* w1, w2: weights of the two BCE terms, loss = w1*BCE(heatmap, T) + w2*BCE(gcn_heatmap, T): (1, 1) for lossDecay == -1, else the
 * (alpha, beta) schedule of /root/reference/misc/losses.py:36-42.
 *
 * hupr_heatmap_bwd: the same two sigmoid backward steps for CALLER-SUPPLIED output gradients g_heat, g_gcn = dL/d(heatmap),
 * dL/d(gcn_heatmap) (float [batch][14][64][64], either may be NULL = zero) — what loss.backward() feeds into the model's outputs when the
 * loss is computed outside the library (/root/reference/tools/run.py:77-78); same output layouts as hupr_heatmap_loss_bwd. */
int hupr_heatmap_bwd(const float* heatmap, const float* gcn_heatmap, const float* g_heat, const float* g_gcn, int batch, int ld,
                     float* d_heat_logits, float* d_gcn_pre, void* stream);
int hupr_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step, const int* step_dev, const float* lr_dev, void* stream);
/* step_dev: optional DEVICE int holding the step count; when given it replaces `step`, so a CUDA-graph replay of the launch keeps
 * the bias corrections advancing.  lr_dev: optional DEVICE float holding the learning rate; when given it replaces `lr`, so the
 * reference's schedule (adjustLR, /root/reference/tools/base.py:66-72) keeps acting on a captured step. */

/* ---------------------------------------------------------------------------------------------
 * Parameter-layout kernels of the training step (csrc/pack.cu).  The optimiser updates one flat fp32 buffer in torch's layouts
 * (what /root/reference/tools/base.py:47 optim.Adam sees); the tensor-core kernels read hi/lo bf16 operand layouts.  These entry points
 * convert inside the captured step, so no framework kernel runs between hupr_adam_step and the next forward.
 *   hupr_pack_conv_weights  w fp32 [cout][cin][taps] (torch conv layout, taps = kd*kh*kw <= 32) ->
 *                             fwd  [taps][cout_total][cin_pad] at rows cout_off..   (hupr_conv_desc.w_*; fwd_lo optional)
 *                             dgrad [taps][cin_pad][cout_total] with the tap order reversed, columns cout_off..  (the flipped/transposed
 *                             filter of the data-gradient convolution; dgrad_* optional).  Padding cells are never written: zero the
 *                             buffers once.  cout, cin, cout_off, cin_pad, cout_total must be even.
 *   hupr_unpack_wgrad       acc fp32 [taps][cin_pad][cout_total] (hupr_conv_wgrad's result) -> dst fp32 [cout][cin][taps], output channels
 *                             cout_off .. cout_off+cout  (the parameter's .grad in torch layout)
 *   hupr_reduce_f64         dst[g] = (float) sum_{i<n} src[g*n + i]   (double partial sums of hupr_channel_sums / hupr_mnet_bwd -> fp32 gradients)
 *   hupr_broadcast_f32      dst[0..n) = *src   (nn.PReLU's single slope -> the per-channel slope array of the conv epilogue)
 *   hupr_gcn_bias_rows      bias fp32 [1024][14] -> bf16 split rows [(b,j)][1024] = bias[q][j] (zero rows past batch*14): the residual operand
 *                             of the transposed-layout GCN GEMMs (/root/reference/models/gcn_networks.py:23-29)
 *   hupr_bump_i32           ++*counter (the device-side Adam step count of hupr_adam_step.step_dev)
 *   hupr_memset_zero        cudaMemsetAsync(dst, 0, bytes) on `stream` (a memset node under CUDA-graph capture) */
int hupr_pack_conv_weights(const float* w, int cout, int cin, int taps, void* fwd_hi, void* fwd_lo, int cout_total, int cin_pad, int cout_off,
                           void* dgrad_hi, void* dgrad_lo, void* stream);
int hupr_unpack_wgrad(const float* acc, int taps, int cin_pad, int cout_total, int cout_off, float* dst, int cout, int cin, void* stream);
/* Batched forms: ONE launch (un)packs every parameter tensor of the model.  `jobs` and `block_job` are DEVICE arrays built once by the
 * caller: jobs[j] describes tensor j with the arguments of the single-tensor entry points; the grid has sum_j blocks_x * ceil(cin/16)
 * CTAs, block_job[b] = the job of CTA b, jobs[j].block_begin = its first CTA.  max_taps = the largest `taps` of the table. */
typedef struct hupr_pack_job {
    const float* w;              /* pack: fp32 source [cout][cin][taps];  unpack: fp32 accumulator [taps][cin_pad][cout_total] */
    void* a_hi; void* a_lo;      /* pack: forward operand planes;         unpack: a_hi = fp32 destination [cout][cin][taps] */
    void* b_hi; void* b_lo;      /* pack: data-gradient operand planes (may be NULL);  unpack: unused */
    int cout, cin, taps;
    int cout_total, cin_pad, cout_off;
    int blocks_x;                /* ceil(cout / 16) */
    int block_begin;
} hupr_pack_job;
int hupr_pack_conv_weights_multi(const hupr_pack_job* jobs, const int* block_job, int total_blocks, int max_taps, void* stream);
int hupr_unpack_wgrad_multi(const hupr_pack_job* jobs, const int* block_job, int total_blocks, int max_taps, void* stream);
int hupr_reduce_f64(const double* src, int groups, int n, float* dst, void* stream);
int hupr_broadcast_f32(const float* src, float* dst, int n, void* stream);
int hupr_gcn_bias_rows(const float* bias, int batch, int rows, void* rows_hi, void* rows_lo, void* stream);
int hupr_bump_i32(int* counter, void* stream);
int hupr_memset_zero(void* dst, size_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HUPR_B200_H */
