"""Host-side mirror of the reference's preprocessing entry point, backed by the sm_100a cascade kernel.

Mirrors /root/reference/preprocessing/process_iwr1843.py (class ``RadarObject``):
  getadcDataFromDCA1000 (:54-83), generateHeatmap (:106-173), saveRadarData (:180-182),
  processRadarDataHoriVert (:184-196) — same names, argument meaning and on-disk format
  (``<saveDir>/{hori,vert}/%09d.npy`` holding a ``[16,64,64,8]`` complex cube).

Differences that matter to a user switching over:
  * the arithmetic runs on the GPU in complex64 (the reference computes complex128 on the CPU);
    ``cubeDtype`` selects what is written to disk (default complex128 so existing loaders see the same
    dtype, values are the complex64 results widened);
  * the GPU path consumes the DCA1000 int16 words directly (``cascade_i16``): 4x fewer bytes over PCIe
    than a complex64 frame, and the de-interleave is fused into the kernel's load stage;
  * there is no CPU fallback — a missing ``libhupr_b200.so`` or a non-B200 device raises RuntimeError.
"""
import os

import numpy as np
import torch

from .. import _C

NUM_RX = 4
NUM_ADC_SAMPLES = 256
NUM_CHIRPS = 64 * 3
FRAME_WORDS = NUM_CHIRPS * NUM_RX * NUM_ADC_SAMPLES * 2      # int16 words per frame-sensor
CUBE_SHAPE = (16, 64, 64, 8)


def cascade_i16(adc, out=None):
    """Run the fused FFT cascade on device-resident DCA1000 words.

    adc : torch.int16 CUDA tensor, ``[n, FRAME_WORDS]`` (or any shape with n*FRAME_WORDS elements).
    Returns torch.complex64 ``[n, 16, 64, 64, 8]`` on the same device.
    """
    if not (isinstance(adc, torch.Tensor) and adc.is_cuda and adc.dtype == torch.int16):
        raise TypeError("cascade_i16 expects a CUDA int16 tensor of DCA1000 words")
    if adc.numel() % FRAME_WORDS != 0:
        raise ValueError("adc has %d words, not a multiple of %d" % (adc.numel(), FRAME_WORDS))
    adc = adc.contiguous()
    n = adc.numel() // FRAME_WORDS
    if out is None:
        out = torch.empty((n,) + CUBE_SHAPE, dtype=torch.complex64, device=adc.device)
    elif out.dtype != torch.complex64 or out.numel() != n * 16 * 64 * 64 * 8 or not out.is_contiguous():
        raise ValueError("out must be a contiguous complex64 tensor [n,16,64,64,8]")
    from .. import ops
    with torch.cuda.device(adc.device), ops._timed("fft_cascade_i16"):
        _C.check(_C.lib().hupr_fft_cascade_i16(_C.ptr(adc), _C.ptr(out), n, _C.stream_ptr()),
                 "hupr_fft_cascade_i16")
    return out


def frames_to_dca1000(frames):
    """complex ``[..., 4, 192, 256]`` (integer-valued) -> int16 DCA1000 words ``[..., FRAME_WORDS]``."""
    frames = np.asarray(frames)
    lead = frames.shape[:-3]
    stream = np.swapaxes(frames, -3, -2).reshape(lead + (-1, 2))          # [..., chirp*rx*sample/2, 2]
    words = np.empty(stream.shape[:-1] + (4,), dtype=np.int16)
    words[..., 0:2] = np.rint(stream.real)
    words[..., 2:4] = np.rint(stream.imag)
    return words.reshape(lead + (-1,))


class RadarObject(object):
    def __init__(self, numGroup=276, root='HuPR', saveRoot='HuPR', device='cuda', cubeDtype=np.complex128,
                 framesPerLaunch=64):
        self.root = root
        self.saveRoot = saveRoot
        self.sensorType = 'iwr1843'
        self.numADCSamples = NUM_ADC_SAMPLES
        self.numRX = NUM_RX
        self.numLanes = 2
        self.framePerSecond = 10
        self.duration = 60
        self.numFrame = self.framePerSecond * self.duration
        self.numChirp = NUM_CHIRPS
        self.device = torch.device(device)
        self.cubeDtype = cubeDtype
        self.framesPerLaunch = framesPerLaunch
        self.radarDataFileNameGroup = []
        self.saveDirNameGroup = []
        self.initialize(numGroup)

    def initialize(self, numGroup):
        for i in range(1, numGroup + 1):
            base = 'raw_data/%s/%s/single_%d' % (self.sensorType, self.root, i)
            self.radarDataFileNameGroup.append([base + '/hori', base + '/vert'])
            self.saveDirNameGroup.append('../data/%s/single_%d' % (self.saveRoot, i))

    # ---- a-1 ---------------------------------------------------------------------------------
    def readDCA1000Words(self, fileName):
        """Raw int16 words of ``<fileName>/adc_data.bin`` — the GPU ingest format."""
        return np.fromfile(os.path.join(fileName, 'adc_data.bin'), dtype=np.int16)

    def getadcDataFromDCA1000(self, fileName):
        """Same result as the reference (:54-83): complex ``[4, nChirp, 256]``, via the closed form."""
        words = self.readDCA1000Words(fileName).reshape(-1, 4)
        stream = words[:, 0:2].reshape(-1).astype(np.float64) + 1j * words[:, 2:4].reshape(-1)
        return stream.reshape(-1, self.numRX, self.numADCSamples).transpose(1, 0, 2)

    # ---- a-2 .. a-7 ---------------------------------------------------------------------------
    def generateHeatmap(self, frame):
        """complex ``[4,192,256]`` frame -> complex ``[16,64,64,8]`` cube (reference :106-173)."""
        words = torch.from_numpy(frames_to_dca1000(np.asarray(frame))).to(self.device, non_blocking=True)
        cube = cascade_i16(words)[0]
        return cube.cpu().numpy().astype(self.cubeDtype, copy=False)

    def generateHeatmapBatch(self, words_i16):
        """int16 words ``[n, FRAME_WORDS]`` (host ndarray or CUDA tensor) -> CUDA complex64 ``[n,16,64,64,8]``."""
        if isinstance(words_i16, np.ndarray):
            words_i16 = torch.from_numpy(np.ascontiguousarray(words_i16)).to(self.device, non_blocking=True)
        return cascade_i16(words_i16)

    # ---- a-8 ---------------------------------------------------------------------------------
    def saveRadarData(self, matrix, dirName, idxFrame):
        np.save(dirName + ('/%09d' % idxFrame) + '.npy', matrix)

    def processRadarDataHoriVert(self):
        """Process every capture directory; file naming identical to the reference (:184-196)."""
        for names, saveDir in zip(self.radarDataFileNameGroup, self.saveDirNameGroup):
            for sensorDir, sub in zip(names, ('hori', 'vert')):
                words = self.readDCA1000Words(sensorDir)
                nFrames = min(self.numFrame, words.size // FRAME_WORDS)
                words = words[:nFrames * FRAME_WORDS].reshape(nFrames, FRAME_WORDS)
                for f0 in range(0, nFrames, self.framesPerLaunch):
                    cubes = self.generateHeatmapBatch(words[f0:f0 + self.framesPerLaunch]).cpu().numpy()
                    for k in range(cubes.shape[0]):
                        self.saveRadarData(cubes[k].astype(self.cubeDtype, copy=False),
                                           saveDir + '/' + sub, f0 + k)
                print('%s, finished %d frames' % (sensorDir, nFrames), end='\r')


if __name__ == "__main__":
    RadarObject().processRadarDataHoriVert()
