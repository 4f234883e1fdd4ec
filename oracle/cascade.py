"""Oracle (TEST INFRASTRUCTURE, never on the product path): numpy fp64 restatement of the
IWR1843 range -> Doppler -> angle FFT cascade.

Follows /root/reference/preprocessing/process_iwr1843.py:
  * dca1000_to_complex      <- getadcDataFromDCA1000   (:54-83)
  * tdm_demux               <- generateHeatmap TDM loop (:113-120)
  * remove_clutter          <- clutterRemoval           (:85-104, called :123-128)
  * generate_heatmap        <- generateHeatmap          (:106-173) incl. postProcessFFT3D (:48-52)
  * generate_heatmap_looped <- same, but with the reference's per-cell call pattern
                               (16 384 x 12 np.fft.fft + 32 768 fftshift calls, :144-164);
                               this is what the reference costs on a CPU and is what
                               bench.py --impl reference times.

Pinned against the reference itself (executed through oracle/ref_shim.py in the build
container) by tests/golden/cascade_*.npz — see oracle/make_golden.py.
"""
import numpy as np

NUM_RX = 4            # process_iwr1843.py:22
NUM_SAMPLES = 256     # :18
NUM_CHIRPS = 192      # :28  (64 chirp loops x 3 TX)
NUM_LOOPS = 64        # :29
NUM_AZ = 64           # :20  numADCSamples // adcRatio
NUM_ELE = 8           # :21
NUM_RANGE_OUT = 64    # :158 numADCSamples // rate
NUM_DOPPLER_OUT = 16  # :159 idxProcChirp // numGroupChirp
RANGE_HI = 94         # :154 idxADCSpecific = 94, 93, ..., 31
FRAME_I16 = NUM_CHIRPS * NUM_RX * NUM_SAMPLES * 2   # int16 words per frame-sensor (786 432 B)


# ----------------------------------------------------------------------------- a-1
def dca1000_to_complex(raw_i16):
    """int16 DCA1000 2-lane stream -> complex128 [4, nChirp, 256]  (:54-83).

    The file is groups of four int16 ``[I(2n), I(2n+1), Q(2n), Q(2n+1)]``; the complex
    stream index is ``chirp*1024 + rx*256 + sample`` (:73-78 walk it in 1024-sample steps).
    """
    raw = np.asarray(raw_i16, dtype=np.int16).reshape(-1, 4)
    re = raw[:, 0:2].reshape(-1).astype(np.float64)
    im = raw[:, 2:4].reshape(-1).astype(np.float64)
    stream = re + 1j * im
    return stream.reshape(-1, NUM_RX, NUM_SAMPLES).transpose(1, 0, 2)


def complex_to_dca1000(frame):
    """Inverse of :func:`dca1000_to_complex` for integer-valued complex [4, nChirp, 256]."""
    stream = np.asarray(frame).transpose(1, 0, 2).reshape(-1)
    out = np.empty((stream.size // 2, 4), dtype=np.int16)
    out[:, 0:2] = np.rint(stream.real).astype(np.int16).reshape(-1, 2)
    out[:, 2:4] = np.rint(stream.imag).astype(np.int16).reshape(-1, 2)
    return out.reshape(-1)


# ----------------------------------------------------------------------------- a-2
def tdm_demux(frame):
    """[4,192,256] -> horizontal virtual array [8,64,256], elevation-TX array [4,64,256] (:113-120)."""
    frame = np.asarray(frame, dtype=np.complex128)
    hor = np.concatenate((frame[:, 0::3, :], frame[:, 2::3, :]), axis=0)
    ver = frame[:, 1::3, :].copy()
    return hor, ver


# ----------------------------------------------------------------------------- a-3
def remove_clutter(x):
    """Subtract the mean over the 64 chirp loops per (antenna, sample) (:85-104, :123-128)."""
    return x - x.mean(axis=1, keepdims=True)


# ----------------------------------------------------------------------------- a-4 .. a-7
def angle_spectrum(hor_rd, ver_rd):
    """Range-Doppler maps -> 4-D spectrum X[ele 8, az 64, doppler 64, range 256] (:137-151).

    Row e=0 holds the 8 horizontal antennas at az slots 0..7, row e=1 the 4 elevation-TX
    antennas at az slots 2..5.  The 8-point elevation FFT is applied ONLY to az slots
    2..5 (:146-149); slots 0,1,6,7 keep [h,0,...,0].  Then a 64-point azimuth FFT for
    every elevation row (:150-151).
    """
    m = np.zeros((NUM_ELE, NUM_AZ) + hor_rd.shape[1:], dtype=np.complex128)
    m[0, 0:8] = hor_rd
    m[1, 2:6] = ver_rd
    m[:, 2:6] = np.fft.fft(m[:, 2:6], axis=0)
    return np.fft.fft(m, axis=1)


def output_index_maps():
    """The four bit-exact integer index maps of the cascade.

    out[d, r, a, e] = X[ele_src[e], az_src[a], dop_src[d], rng_src[r]]
    derived from :153-171 + postProcessFFT3D :48-52.
    """
    ele_src = (11 - np.arange(NUM_ELE)) % NUM_ELE
    az_src = (95 - np.arange(NUM_AZ)) % NUM_AZ
    dop_src = (np.arange(NUM_DOPPLER_OUT) + 56) % NUM_LOOPS
    rng_src = RANGE_HI - np.arange(NUM_RANGE_OUT)
    return ele_src, az_src, dop_src, rng_src


def generate_heatmap(frame):
    """complex [4,192,256] -> complex128 [16,64,64,8] (doppler, range, azimuth, elevation)."""
    hor, ver = tdm_demux(frame)
    hor = np.fft.fft2(remove_clutter(hor), axes=(1, 2))      # :131-134
    ver = np.fft.fft2(remove_clutter(ver), axes=(1, 2))
    spec = angle_spectrum(hor, ver)
    ele_src, az_src, dop_src, rng_src = output_index_maps()
    out = spec[np.ix_(ele_src, az_src, dop_src, rng_src)]     # [e, a, d, r]
    return np.ascontiguousarray(out.transpose(2, 3, 1, 0))


def generate_heatmap_looped(frame):
    """Same result as :func:`generate_heatmap` with the reference's CPU cost structure:
    one np.fft.fft call per (chirp, sample, column) and one fftshift per (ele, az, range)
    (:144-164).  Used only to time "what the reference costs" on the bench host."""
    hor, ver = tdm_demux(frame)
    hor = remove_clutter(hor)
    ver = remove_clutter(ver)
    for ant in range(hor.shape[0]):
        hor[ant] = np.fft.fft2(hor[ant])
    for ant in range(ver.shape[0]):
        ver[ant] = np.fft.fft2(ver[ant])
    cube = np.zeros((NUM_ELE, NUM_AZ, NUM_LOOPS, NUM_SAMPLES), dtype=np.complex128)
    cube[0, 0:8] = hor
    cube[1, 2:6] = ver
    fft = np.fft.fft
    for c in range(NUM_LOOPS):
        for s in range(NUM_SAMPLES):
            cell = cube[:, :, c, s]
            for z in (2, 3, 4, 5):
                cell[:, z] = fft(cell[:, z])
            for e in range(NUM_ELE):
                cell[e, :] = fft(cell[e, :])
    shifted = np.zeros((NUM_LOOPS, NUM_RANGE_OUT, NUM_AZ, NUM_ELE), dtype=np.complex128)
    shift = np.fft.fftshift
    for e in range(NUM_ELE):
        for a in range(NUM_AZ):
            for r in range(NUM_RANGE_OUT):
                shifted[:, r, a, e] = shift(cube[e, a, :, RANGE_HI - r], axes=(0))
    lo = NUM_LOOPS // 2 - NUM_DOPPLER_OUT // 2
    out = np.zeros((NUM_DOPPLER_OUT, NUM_RANGE_OUT, NUM_AZ, NUM_ELE), dtype=np.complex128)
    for d in range(NUM_DOPPLER_OUT):
        plane = np.fft.fftshift(shifted[lo + d].transpose(1, 2, 0), axes=(0, 1))   # [a, e, r]
        out[d] = plane.transpose(2, 0, 1)[:, ::-1, ::-1]
    return out


# ----------------------------------------------------------------------------- synthetic input
def synth_frame(frame_idx, sensor=0, n_targets=3):
    """Config-1 synthetic IWR1843 frame (SURVEY.md §8 d): 12-bit-ish ADC noise plus
    ``n_targets`` point targets.  Returns integer-valued complex128 [4,192,256]."""
    rng = np.random.default_rng(1000 * sensor + frame_idx)
    re = rng.integers(-2048, 2048, (NUM_RX, NUM_CHIRPS, NUM_SAMPLES)).astype(np.float64)
    im = rng.integers(-2048, 2048, (NUM_RX, NUM_CHIRPS, NUM_SAMPLES)).astype(np.float64)
    sig = np.zeros((NUM_RX, NUM_CHIRPS, NUM_SAMPLES), dtype=np.complex128)
    rx = np.arange(NUM_RX)[:, None, None]
    chirp = np.arange(NUM_CHIRPS)[None, :, None]
    samp = np.arange(NUM_SAMPLES)[None, None, :]
    for _ in range(n_targets):
        rbin = rng.integers(31, 95)
        dop = rng.choice([d for d in range(-7, 8) if d != 0])
        amp = rng.uniform(200.0, 1000.0)
        az = rng.uniform(-0.8, 0.8) * np.pi
        el = rng.uniform(-0.8, 0.8) * np.pi
        tx = chirp % 3                      # 0: az TX0, 1: elevation TX, 2: az TX (+4 antennas)
        loop = chirp // 3
        ant_az = rx + 4 * (tx == 2) + 2 * (tx == 1)
        phase = (2 * np.pi * rbin * samp / NUM_SAMPLES
                 + 2 * np.pi * dop * loop / NUM_LOOPS
                 + az * ant_az + el * (tx == 1))
        sig += amp * np.exp(1j * phase)
    frame = np.clip(np.rint(re + sig.real), -32768, 32767) + 1j * np.clip(np.rint(im + sig.imag), -32768, 32767)
    return frame


def synth_raw_i16(frame_idx, sensor=0):
    """DCA1000-format int16 words for :func:`synth_frame` (length FRAME_I16)."""
    return complex_to_dca1000(synth_frame(frame_idx, sensor))
