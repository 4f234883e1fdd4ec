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

#ifdef __cplusplus
}
#endif
#endif /* HUPR_B200_H */
