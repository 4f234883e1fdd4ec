// IWR1843 range -> Doppler -> angle FFT cascade, one fused sm_100a kernel.
//
// Replaces RadarObject.generateHeatmap (+ the DCA1000 de-interleave feeding it):
//   /root/reference/preprocessing/process_iwr1843.py:54-83   int16 2-lane IQ de-interleave
//   /root/reference/preprocessing/process_iwr1843.py:106-173 TDM demux, clutter removal, fft2,
//                                                            elevation/azimuth FFTs, crop/shift/flip
//
// Work decomposition:
//   * one 2-CTA thread-block cluster per frame-sensor (persistent, grid = #SMs);
//   * each CTA range/Doppler-transforms 6 of the 12 virtual antennas:
//       - chirp rows (1 KiB of int16 IQ each) are staged by the TMA engine (cp.async.bulk + mbarrier,
//         double buffered, prefetched two passes ahead, also across frame-sensor boundaries);
//       - 256-pt range FFT = 16 x 16 four-step: radix-16 in registers, one swizzled shared-memory
//         exchange, pruned radix-16 (only range bins 31..94 are ever produced);
//       - clutter removal + 64-pt Doppler FFT pruned to the 16 surviving bins: radix-8 in registers,
//         then a warp-shuffle reduce-scatter over 8 lanes;
//   * the range-Doppler planes are exchanged through distributed shared memory so that CTA k owns all
//     12 antennas for Doppler half k; cluster barriers order the exchange;
//   * angle stage: one warp per (doppler, range) cell, 12 inputs -> 64 az x 8 ele outputs, written as
//     512-byte fully coalesced warp stores straight to the [16,64,64,8] cube (index maps of
//     SURVEY.md §8 a-7 are folded into the lane mapping).
#include "common.cuh"

namespace hupr {

constexpr int kCascadeThreads = 512;
constexpr int kRowBytes = 1024;                 // 256 complex int16 samples
constexpr int kPassRows = 32;
constexpr int kPassBytes = kPassRows * kRowBytes;
constexpr int kPassesPerFs = 12;                // 6 antennas per CTA x 2 chirp halves
constexpr int kFsRows = 768;                    // 192 chirps x 4 RX
constexpr int kCubeElems = 16 * 64 * 64 * 8;    // float2 elements per frame-sensor

constexpr int SM_RD = 0;                                   // [8 dl][64 r][12 v] float2
constexpr int SM_RANGE = SM_RD + 8 * 64 * 12 * 8;          // [64 r][64 chirp] float2 (xor swizzled)
constexpr int SM_EXCH = SM_RANGE + 64 * 64 * 8;            // [32 rows][16 k1][16 n2] float2 (swizzled)
constexpr int SM_RAW = SM_EXCH + kPassRows * 2048;         // 2 x [32 rows][1 KiB]
constexpr int SM_TW1 = SM_RAW + 2 * kPassBytes;            // [16 k1][16 n2] float2
constexpr int SM_TWD = SM_TW1 + 2048;                      // [8 rho][8 m2] float2
constexpr int SM_BAR = SM_TWD + 512;                       // 2 mbarriers
constexpr int SM_CASCADE_TOTAL = SM_BAR + 64;

__device__ __forceinline__ float2 unpack_iq(uint2 w, uint32_t sel) {
    int i, q;
    asm("prmt.b32 %0, %1, 0, %2;" : "=r"(i) : "r"(w.x), "r"(sel));
    asm("prmt.b32 %0, %1, 0, %2;" : "=r"(q) : "r"(w.y), "r"(sel));
    return make_float2(__int2float_rn(i), __int2float_rn(q));
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kCascadeThreads, 1)
cascade_kernel(const int16_t* __restrict__ adc, float2* __restrict__ cube, int n_fs) {
    extern __shared__ __align__(128) uint8_t smem[];
    float2* sRD = reinterpret_cast<float2*>(smem + SM_RD);
    float2* sRange = reinterpret_cast<float2*>(smem + SM_RANGE);
    uint8_t* sExch = smem + SM_EXCH;
    uint8_t* sRaw = smem + SM_RAW;
    float2* sTw1 = reinterpret_cast<float2*>(smem + SM_TW1);
    float2* sTwd = reinterpret_cast<float2*>(smem + SM_TWD);
    uint64_t* sFull = reinterpret_cast<uint64_t*>(smem + SM_BAR);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1;
    const int n_clusters = gridDim.x >> 1;

    // ---- one-time tables ------------------------------------------------------------------
    if (tid < 256) {
        const int k1 = tid >> 4, n2 = tid & 15;
        const int b = (k1 == 15) ? 1 : 2;          // output-bin shift folded into the step-1 twiddle
        sTw1[tid] = twiddle((n2 * (k1 + 16 * b)) & 255, 256);
    }
    if (tid < 64) sTwd[tid] = twiddle(((tid >> 3) * (tid & 7)) & 63, 64);
    if (tid == 0) {
        mbar_init(&sFull[0], 1);
        mbar_init(&sFull[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    cluster_arrive();
    cluster_wait();      // peer CTA is resident and initialised
    cluster_arrive();    // "my RD buffer may be overwritten" token for iteration 0

    const int n_my = (cluster_id < n_fs) ? (n_fs - cluster_id + n_clusters - 1) / n_clusters : 0;
    const int total_passes = n_my * kPassesPerFs;

    auto issue_pass = [&](int g) {   // executed by warp 0; lane j fetches chirp row j of the pass
        const int it = g / kPassesPerFs, p = g - it * kPassesPerFs;
        const int v = 6 * (int)rank + (p >> 1), h = p & 1;
        const int tx = (v < 4) ? 0 : ((v < 8) ? 2 : 1);
        const int rx = v & 3;
        const size_t fs = (size_t)cluster_id + (size_t)it * n_clusters;
        const int buf = g & 1;
        if (lane == 0) mbar_expect_tx(&sFull[buf], kPassBytes);
        __syncwarp();
        const int chirp = 3 * (32 * h + lane) + tx;
        const int16_t* src = adc + (fs * kFsRows + (size_t)(chirp * 4 + rx)) * 512;
        bulk_g2s(sRaw + buf * kPassBytes + lane * kRowBytes, src, kRowBytes, &sFull[buf]);
    };
    if (warp == 0) {
        if (total_passes > 0) issue_pass(0);
        if (total_passes > 1) issue_pass(1);
    }

    // ---- per-lane constants of the angle stage --------------------------------------------
    const int q4 = lane & 3;
    const int a0 = 7 - (lane >> 2);
    const int ep0 = (11 - 2 * q4) & 7;   // source elevation bin of output e = 2*q4
    const int ep1 = (10 - 2 * q4) & 7;   // source elevation bin of output e = 2*q4 + 1
    const float e0mask = (ep1 == 0) ? 1.0f : 0.0f;   // ep0 is never 0

    for (int g = 0; g < total_passes; ++g) {
        const int it = g / kPassesPerFs, p = g - it * kPassesPerFs;
        const int v = 6 * (int)rank + (p >> 1), h = p & 1;
        const int buf = g & 1;
        mbar_wait(&sFull[buf], (uint32_t)((g >> 1) & 1));

        // ================= range FFT step 1: 16-pt DFT over n1 of x[16*n1 + n2] =================
        {
            const int j = tid >> 4, n2 = tid & 15;
            const uint2* rowp = reinterpret_cast<const uint2*>(sRaw + buf * kPassBytes + j * kRowBytes);
            const uint32_t sel = (n2 & 1) ? 0xBB32u : 0x9910u;   // pick + sign-extend one int16 half
            float2 x[16];
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) x[n1] = unpack_iq(rowp[8 * n1 + (n2 >> 1)], sel);
            fft16(x);
            uint8_t* erow = sExch + j * 2048 + ((n2 & 1) << 3);
            const int c = n2 >> 1;
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) {
                const float2 t = cmul(x[k1], sTw1[k1 * 16 + n2]);
                *reinterpret_cast<float2*>(erow + k1 * 128 + ((c ^ (k1 & 7)) << 4)) = t;
            }
        }
        __syncthreads();   // raw[buf] consumed, exchange written
        if (warp == 0 && g + 2 < total_passes) issue_pass(g + 2);

        // ================= range FFT step 2: pruned 16-pt DFT over n2 (4 outputs) ===============
        {
            const int j = tid >> 4, k1 = tid & 15;
            const uint8_t* eline = sExch + j * 2048 + k1 * 128;
            float2 y[16];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 q = *reinterpret_cast<const float4*>(eline + ((c ^ (k1 & 7)) << 4));
                y[2 * c] = make_float2(q.x, q.y);
                y[2 * c + 1] = make_float2(q.z, q.w);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) fft4(y[t], y[4 + t], y[8 + t], y[12 + t]);   // P_t[u] -> y[4u+t]
            const float r2 = 0.70710678118654752440f;
            const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
            float2 X[4];
            X[0] = cadd(cadd(y[0], y[1]), cadd(y[2], y[3]));
            X[1] = cfma(make_float2(c1, -s1), y[5], y[4]);
            X[1] = cfma(make_float2(r2, -r2), y[6], X[1]);
            X[1] = cfma(make_float2(s1, -c1), y[7], X[1]);
            X[2] = cfma(make_float2(r2, -r2), y[9], y[8]);
            X[2] = cadd(X[2], mul_neg_i(y[10]));
            X[2] = cfma(make_float2(-r2, -r2), y[11], X[2]);
            X[3] = cfma(make_float2(s1, -c1), y[13], y[12]);
            X[3] = cfma(make_float2(-r2, -r2), y[14], X[3]);
            X[3] = cfma(make_float2(-c1, s1), y[15], X[3]);
            const int b = (k1 == 15) ? 1 : 2;
            const int m = 32 * h + j;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = 94 - k1 - 16 * (b + u);
                const int f = ((r & 1) << 3) | ((r >> 1) & 7);
                sRange[r * 64 + (m ^ f)] = X[u];
            }
        }
        __syncthreads();   // exchange free, range buffer written

        // ================= clutter removal + pruned 64-pt Doppler FFT ===========================
        if (h == 1) {
            if (p == 1) cluster_wait();   // peer finished reading its RD buffer (previous frame-sensor)
            const int r = tid >> 3, m2 = tid & 7;
            const int f = ((r & 1) << 3) | ((r >> 1) & 7);
            float2 x[8];
#pragma unroll
            for (int m1 = 0; m1 < 8; ++m1) x[m1] = sRange[r * 64 + ((8 * m1 + m2) ^ f)];
            float2 s = cadd(cadd(cadd(x[0], x[1]), cadd(x[2], x[3])), cadd(cadd(x[4], x[5]), cadd(x[6], x[7])));
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
                s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
            }
            const float2 mean = make_float2(s.x * (1.0f / 64.0f), s.y * (1.0f / 64.0f));
#pragma unroll
            for (int m1 = 0; m1 < 8; ++m1) x[m1] = csub(x[m1], mean);
            fft8(x);                                   // Q[rho], rho = d' mod 8
            const float2 w7 = twiddle((7 * m2) & 7, 8);  // w8^(7*m2): selects d' = 56 + rho
            float2 U[8], W[8];
#pragma unroll
            for (int rho = 0; rho < 8; ++rho) {
                U[rho] = cmul(x[rho], sTwd[rho * 8 + m2]);   // d' = rho      -> out d = 8 + rho
                W[rho] = cmul(U[rho], w7);                   // d' = 56 + rho -> out d = rho
            }
            // reduce-scatter the 16 partial sums over the 8 lanes of this column
            const bool b2 = (m2 & 4) != 0, b1 = (m2 & 2) != 0, b0 = (m2 & 1) != 0;
            float2 k8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float2 keep = b2 ? W[i] : U[i];
                const float2 send = b2 ? U[i] : W[i];
                k8[i] = make_float2(keep.x + __shfl_xor_sync(0xffffffffu, send.x, 4),
                                    keep.y + __shfl_xor_sync(0xffffffffu, send.y, 4));
            }
            float2 k4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 keep = b1 ? k8[4 + i] : k8[i];
                const float2 send = b1 ? k8[i] : k8[4 + i];
                k4[i] = make_float2(keep.x + __shfl_xor_sync(0xffffffffu, send.x, 2),
                                    keep.y + __shfl_xor_sync(0xffffffffu, send.y, 2));
            }
            float2 k2[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float2 keep = b0 ? k4[2 + i] : k4[i];
                const float2 send = b0 ? k4[i] : k4[2 + i];
                k2[i] = make_float2(keep.x + __shfl_xor_sync(0xffffffffu, send.x, 1),
                                    keep.y + __shfl_xor_sync(0xffffffffu, send.y, 1));
            }
            const uint32_t dst_rank = b2 ? 0u : 1u;
            const int rho0 = (b1 ? 4 : 0) + (b0 ? 2 : 0);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                float2* dst = sRD + ((rho0 + i) * 64 + r) * 12 + v;
                if (dst_rank == rank) *dst = k2[i];
                else st_cluster_f2(mapa(smem_u32(dst), dst_rank), k2[i]);
            }
        }

        // ================= angle stage (after the last antenna of this frame-sensor) ============
        if (p == kPassesPerFs - 1) {
            cluster_arrive();
            cluster_wait();    // both halves of the range-Doppler exchange have landed
            float2 tw[8], c20[4], c21[4];
#pragma unroll
            for (int z = 0; z < 8; ++z) tw[z] = twiddle((a0 * z) & 63, 64);
            {
                const float2 we0 = twiddle(ep0, 8), we1 = twiddle(ep1, 8);
#pragma unroll
                for (int z = 0; z < 4; ++z) {
                    c20[z] = cmul(we0, tw[z + 2]);
                    c21[z] = cmul(we1, tw[z + 2]);
                }
            }
            const size_t fs = (size_t)cluster_id + (size_t)it * n_clusters;
            for (int cell = warp; cell < 512; cell += kCascadeThreads / 32) {
                const float4* in = reinterpret_cast<const float4*>(sRD + cell * 12);
                float2 H[8], V[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 q = in[i];
                    H[2 * i] = make_float2(q.x, q.y);
                    H[2 * i + 1] = make_float2(q.z, q.w);
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float4 q = in[4 + i];
                    V[2 * i] = make_float2(q.x, q.y);
                    V[2 * i + 1] = make_float2(q.z, q.w);
                }
                float2 x0[8], x1[8];
                x0[0] = x0[1] = x0[6] = x0[7] = make_float2(0.f, 0.f);
                x1[0] = make_float2(H[0].x * e0mask, H[0].y * e0mask);          // tw[0] == 1
                {
                    const float2 t1 = cmul(H[1], tw[1]), t6 = cmul(H[6], tw[6]), t7 = cmul(H[7], tw[7]);
                    x1[1] = make_float2(t1.x * e0mask, t1.y * e0mask);
                    x1[6] = make_float2(t6.x * e0mask, t6.y * e0mask);
                    x1[7] = make_float2(t7.x * e0mask, t7.y * e0mask);
                }
#pragma unroll
                for (int z = 0; z < 4; ++z) {
                    const float2 hz = cmul(H[z + 2], tw[z + 2]);
                    x0[z + 2] = cfma(c20[z], V[z], hz);
                    x1[z + 2] = cfma(c21[z], V[z], hz);
                }
                fft8(x0);
                fft8(x1);
                const int dl = cell >> 6, r = cell & 63;
                float4* outp = reinterpret_cast<float4*>(
                                   cube + ((fs * 16 + (size_t)(8 * rank + dl)) * 64 + r) * 512) + lane;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int a1 = (11 - i) & 7;
                    st_global_cs_f4(outp + 32 * i, make_float4(x0[a1].x, x0[a1].y, x1[a1].x, x1[a1].y));
                }
            }
            cluster_arrive();   // my RD buffer may be overwritten by the next frame-sensor
        }
    }
    cluster_wait();   // consume the last token; peer is done with my shared memory
}

}  // namespace hupr

extern "C" int hupr_fft_cascade_i16(const int16_t* adc, void* cube, int n_frame_sensors, void* stream) {
    using namespace hupr;
    if (n_frame_sensors < 0) return HUPR_ERR_BAD_ARG;
    if (n_frame_sensors == 0) return HUPR_OK;
    if (adc == nullptr || cube == nullptr) return HUPR_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(adc) & 15) || (reinterpret_cast<uintptr_t>(cube) & 15)) return HUPR_ERR_ALIGNMENT;
    if (int arch_rc = device_check_sm100()) return arch_rc;
    const int num_sms = device_sm_count();
    static bool configured[kMaxDevices] = {};
    if (int crc = ensure_smem_optin(cascade_kernel, SM_CASCADE_TOTAL, configured)) return crc;
    int clusters = num_sms / 2;
    if (clusters > n_frame_sensors) clusters = n_frame_sensors;
    cascade_kernel<<<2 * clusters, kCascadeThreads, SM_CASCADE_TOTAL, static_cast<cudaStream_t>(stream)>>>(
        adc, static_cast<float2*>(cube), n_frame_sensors);
    note_launches(1);
    return (cudaGetLastError() == cudaSuccess) ? HUPR_OK : HUPR_ERR_CUDA;
}
