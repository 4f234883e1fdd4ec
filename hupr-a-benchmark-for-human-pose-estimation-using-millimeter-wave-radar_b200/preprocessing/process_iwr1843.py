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
  * there is no CPU fallback — a missing ``libhupr_b200.so`` or a non-B200 device raises RuntimeError;
  * ``processRadarDataHoriVert`` is a three-stage pipeline: the capture file is read through a memory map into
    PINNED staging buffers, uploaded on a copy stream while the previous chunk is in the cascade, and the results return through pinned
    buffers on a third stream while a writer thread saves the files of the chunk before;
  * ``cacheFormat='planes'`` stores what the loader actually consumes — the 8 kept Doppler rows, real and imaginary
    parts standardised per (row, part, elevation) plane, float32 ``[8,2,64,64,8]`` = 2 MiB per frame instead of the 8 MiB complex128
    cube — computed by the same kernel the loader would run (``hupr_window_normalize``), so a loader item built from either cache is
    bit-identical; ``hupr_b200.datasets.cubecache.convert`` converts an existing cube cache.
"""
import concurrent.futures
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
                 framesPerLaunch=64, cacheFormat='cube'):
        if cacheFormat not in ('cube', 'planes'):
            raise ValueError("cacheFormat must be 'cube' (reference .npy cubes) or 'planes' (compact standardised planes)")
        self.cacheFormat = cacheFormat
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

    # ---- f-3: pipelined ingest of one capture file ----------------------------------------------
    def _stage_buffers(self):
        L, dev = self.framesPerLaunch, self.device
        planes = self.cacheFormat == 'planes'
        out_shape = (L, 8, 2, 64, 64, 8) if planes else (L,) + CUBE_SHAPE
        out_dtype = torch.float32 if planes else torch.complex64
        return {"h_in": torch.empty((L, FRAME_WORDS), dtype=torch.int16).pin_memory(),
                "d_in": torch.empty((L, FRAME_WORDS), dtype=torch.int16, device=dev),
                "d_cube": torch.empty((L,) + CUBE_SHAPE, dtype=torch.complex64, device=dev),
                "d_out": torch.empty(out_shape, dtype=out_dtype, device=dev) if planes else None,
                "h_out": torch.empty(out_shape, dtype=out_dtype).pin_memory(),
                "slots": torch.arange(L, dtype=torch.int32, device=dev),
                "uploaded": torch.cuda.Event(), "computed": torch.cuda.Event(), "returned": torch.cuda.Event(), "pending": None}

    def _save_chunk(self, stage, n, saveDir, f0):
        stage["returned"].synchronize()
        host = stage["h_out"][:n].numpy()
        for k in range(n):
            frame = host[k] if self.cacheFormat == 'planes' else host[k].astype(self.cubeDtype, copy=False)
            self.saveRadarData(frame, saveDir, f0 + k)

    def processCapture(self, words, saveDir, nFrames=None):
        """int16 words of one capture (ndarray or np.memmap, ``nFrames * FRAME_WORDS`` long) -> ``<saveDir>/%09d.npy`` per frame.
        Upload (copy stream), cascade (+ standardisation for the planes cache), read-back (third stream) and file writing (thread) of
        consecutive chunks overlap; two stages of pinned / device buffers are re-used for the whole capture."""
        from .. import ops
        words = np.asarray(words) if not isinstance(words, np.memmap) else words
        total = words.size // FRAME_WORDS if nFrames is None else nFrames
        words = words[:total * FRAME_WORDS].reshape(total, FRAME_WORDS)
        if not hasattr(self, "_stages"):
            with torch.cuda.device(self.device):
                self._stages = [self._stage_buffers(), self._stage_buffers()]
                self._copy_in, self._copy_out = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
                self._writer = concurrent.futures.ThreadPoolExecutor(max_workers=1)
        L = self.framesPerLaunch
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream(self.device)
            for c, f0 in enumerate(range(0, total, L)):
                st = self._stages[c & 1]
                n = min(L, total - f0)
                if st["pending"] is not None:
                    st["pending"].result()                       # the files of this stage's previous chunk are on disk: buffers are free
                np.copyto(st["h_in"][:n].numpy(), words[f0:f0 + n])
                with torch.cuda.stream(self._copy_in):
                    st["d_in"][:n].copy_(st["h_in"][:n], non_blocking=True)
                    st["uploaded"].record(self._copy_in)
                main.wait_event(st["uploaded"])
                cascade_i16(st["d_in"][:n], st["d_cube"][:n])
                if self.cacheFormat == 'planes':
                    ops.window_normalize(st["d_cube"], st["slots"][:n], st["d_out"][:n])
                st["computed"].record(main)
                with torch.cuda.stream(self._copy_out):
                    self._copy_out.wait_event(st["computed"])
                    src = st["d_out"] if self.cacheFormat == 'planes' else st["d_cube"]
                    st["h_out"][:n].copy_(src[:n], non_blocking=True)
                    st["returned"].record(self._copy_out)
                st["pending"] = self._writer.submit(self._save_chunk, st, n, saveDir, f0)
            for st in self._stages:
                if st["pending"] is not None:
                    st["pending"].result()
                    st["pending"] = None
        return total

    def processRadarDataHoriVert(self):
        """Process every capture directory; file naming identical to the reference (:184-196)."""
        for names, saveDir in zip(self.radarDataFileNameGroup, self.saveDirNameGroup):
            for sensorDir, sub in zip(names, ('hori', 'vert')):
                words = np.memmap(os.path.join(sensorDir, 'adc_data.bin'), dtype=np.int16, mode='r')
                nFrames = min(self.numFrame, words.size // FRAME_WORDS)
                os.makedirs(saveDir + '/' + sub, exist_ok=True)
                self.processCapture(words, saveDir + '/' + sub, nFrames)
                print('%s, finished %d frames' % (sensorDir, nFrames), end='\r')


if __name__ == "__main__":
    RadarObject().processRadarDataHoriVert()
