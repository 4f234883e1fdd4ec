"""End-to-end radar -> pose stream on one B200: DCA1000 int16 words -> FFT cascade -> window/standardise -> HuPRNet ->
argmax keypoints, with every intermediate resident in HBM and the whole step replayed as one CUDA graph.

This is the composition the reference spreads over two programs and the file system
(/root/reference/preprocessing/process_iwr1843.py:184-196 -> .npy files -> datasets/dataset.py:120-159 ->
tools/run.py:35-58); the arithmetic of each stage is unchanged, only the hand-offs stay on the device.
"""
import torch

from . import ops
from .ops import SplitTensor
from .preprocessing.process_iwr1843 import FRAME_WORDS, cascade_i16

GROUP = 8


def window_slots(n_windows, n_frames, first_cube=0, group=GROUP):
    """Cube index per window slot for a contiguous stream: window b covers frames b .. b+group-1 (centre frame b + group/2,
    i.e. the reference window [index-4, index+3] of dataset.py:126-138 away from the capture boundaries)."""
    if n_windows + group - 1 > n_frames:
        raise ValueError("%d windows need %d frames, got %d" % (n_windows, n_windows + group - 1, n_frames))
    b = torch.arange(n_windows, dtype=torch.int32).view(-1, 1)
    g = torch.arange(group, dtype=torch.int32).view(1, -1)
    return (first_cube + b + g).reshape(-1).contiguous()


class RadarPoseStream(object):
    """Fixed-shape streaming step: ``n_windows`` poses from ``n_windows + 7`` consecutive radar frames (hori + vert)."""

    def __init__(self, model, n_windows, device="cuda", use_graph=True, per_frame=True):
        """``per_frame=True`` (default): standardisation + MNet run once per frame-sensor and the windows are overlapping strided
        views of the per-frame features (same arithmetic, 8x less pre-encoder work).  ``per_frame=False``
        materialises the reference's ``[B,8,8,2,64,64,8]`` network inputs and calls ``model(hori, vert)``."""
        self.model = model
        self.n_windows = n_windows
        self.n_frames = n_windows + GROUP - 1
        self.device = torch.device(device)
        dev = self.device
        nfs = 2 * self.n_frames
        self.adc = torch.zeros((nfs, FRAME_WORDS), dtype=torch.int16, device=dev)          # [hori frames | vert frames]
        self.cubes = torch.empty((nfs, 16, 64, 64, 8), dtype=torch.complex64, device=dev)
        self.slots_hori = window_slots(n_windows, self.n_frames, 0).to(dev)
        self.slots_vert = window_slots(n_windows, self.n_frames, self.n_frames).to(dev)
        self.per_frame = per_frame
        if per_frame:
            split = model.split
            self.stats = torch.empty(nfs * 256, dtype=torch.float32, device=dev)
            self.feat_hori = SplitTensor.empty((self.n_frames, 64, 64, 32), dev, split)
            self.feat_vert = SplitTensor.empty((self.n_frames, 64, 64, 32), dev, split)
            frame = 64 * 64 * 32
            view = lambda t: None if t is None else t.as_strided((n_windows, GROUP, 64, 64, 32), (frame, frame, 64 * 32, 32, 1))
            self.win_hori = SplitTensor(view(self.feat_hori.hi), view(self.feat_hori.lo))
            self.win_vert = SplitTensor(view(self.feat_vert.hi), view(self.feat_vert.lo))
        else:
            self.vrdae_hori = torch.empty((n_windows, GROUP, 8, 2, 64, 64, 8), dtype=torch.float32, device=dev)
            self.vrdae_vert = torch.empty_like(self.vrdae_hori)
        self.keypoints = torch.empty((n_windows, 14, 2), dtype=torch.float32, device=dev)
        self.maxvals = torch.empty((n_windows, 14), dtype=torch.float32, device=dev)
        self.heatmap = None
        self.gcn_heatmap = None
        self.launches_per_step = 0
        self.graph = None
        self.use_graph = use_graph

    def _step_eager(self):
        cascade_i16(self.adc, self.cubes)
        if self.per_frame:
            ops.plane_stats(self.cubes, self.stats)
            wh, bh, wv, bv = self.model.chirp_weights()
            ops.frame_features(self.cubes, self.stats, 0, self.n_frames, wh, bh, self.feat_hori)
            ops.frame_features(self.cubes, self.stats, self.n_frames, self.n_frames, wv, bv, self.feat_vert)
            heat, gcn = self.model.forward_features(self.win_hori, self.win_vert)
            heat, gcn = heat.unsqueeze(2), gcn.unsqueeze(1)
        else:
            ops.window_normalize(self.cubes, self.slots_hori, self.vrdae_hori)
            ops.window_normalize(self.cubes, self.slots_vert, self.vrdae_vert)
            heat, gcn = self.model(self.vrdae_hori, self.vrdae_vert)
        self.heatmap, self.gcn_heatmap = heat, gcn
        ops.keypoints_argmax(gcn.view(self.n_windows, 14, 64, 64), self.keypoints, self.maxvals)

    def prepare(self):
        """Warm up (lazy weight packing, function attributes) and capture the step into a CUDA graph."""
        with torch.cuda.device(self.device):
            self._step_eager()
            before = ops.launch_count()
            self._step_eager()
            self.launches_per_step = ops.launch_count() - before
            torch.cuda.synchronize()
            if self.use_graph:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self._step_eager()
                self.graph = graph
                torch.cuda.synchronize()
        return self

    def step(self):
        """Run one step on the current contents of ``self.adc``; results land in ``self.keypoints`` (device)."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step_eager()
        return self.keypoints

    # ---- pipelined ingest: the H2D copy of the NEXT frames runs on a copy stream while the current step computes ----------------
    def prefetch(self, adc_hori, adc_vert):
        """Start the asynchronous upload of the next step's DCA1000 words (pinned host tensors) into a staging buffer."""
        if not hasattr(self, "_staging"):
            self._staging = torch.empty_like(self.adc)
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._uploaded = torch.cuda.Event()
            self._consumed = torch.cuda.Event()
            self._consumed.record(torch.cuda.current_stream(self.device))
        n = self.n_frames
        self._copy_stream.wait_event(self._consumed)            # the previous staging contents have been moved into self.adc
        with torch.cuda.stream(self._copy_stream):
            self._staging[:n].copy_(adc_hori.view(n, FRAME_WORDS), non_blocking=True)
            self._staging[n:].copy_(adc_vert.view(n, FRAME_WORDS), non_blocking=True)
            self._uploaded.record(self._copy_stream)

    def step_prefetched(self):
        """Run one step on the frames uploaded by the last ``prefetch`` call; returns the device keypoints ``[n_windows, 14, 2]``."""
        main = torch.cuda.current_stream(self.device)
        main.wait_event(self._uploaded)
        self.adc.copy_(self._staging, non_blocking=True)         # device-to-device, ~20 us
        self._consumed.record(main)
        return self.step()

    def __call__(self, adc_hori, adc_vert):
        """adc_hori / adc_vert: int16 ``[n_frames, FRAME_WORDS]`` (host pinned or device) -> keypoints ``[n_windows, 14, 2]`` (device)."""
        n = self.n_frames
        self.adc[:n].copy_(adc_hori.view(n, FRAME_WORDS), non_blocking=True)
        self.adc[n:].copy_(adc_vert.view(n, FRAME_WORDS), non_blocking=True)
        return self.step()
