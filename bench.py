#!/usr/bin/env python
"""bench.py — headline benchmark of the hupr_b200 hot path (contract: task statement / DESIGN.md §bench).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload e2e|forward-b1|cascade] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): radar frames/s end-to-end (preproc + forward); one radar frame = one hori + one vert IWR1843
capture (2 x 786 432 B of int16 IQ) and yields one 14-keypoint pose.

Workloads:
  e2e         BASELINE.json configs[2] (default): 39 consecutive synthetic frames x 2 sensors -> FFT cascade ->
              window/standardise -> MSCSA-PRGCN forward (batch 32) -> argmax keypoints [32,14,2]
  forward-b1  configs[1]: forward-only inference, batch 1 (CUDA-graph replay latency)
  cascade     configs[4]: FFT-cascade throughput on 2048 frame-sensors per step
  cascade-sweep  configs[4] in full: N = 1k, 4k, 16k, 64k, 256k, 1M radar frames (sharded over the ranks) through the cascade in
              resident chunks of 1024 frames; one JSON line with the per-N table ("sweep")
  train       configs[3] (per-GPU shard): training step, batch --batch per GPU, one NCCL gradient all-reduce when N > 1
A "step" is one pass of the workload over one batch of synthetic input resident in HBM.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "radar_frames_per_sec"
UNIT = "frames/s"
FS_IN_BYTES = 786432            # int16 ADC words per frame-sensor
FS_OUT_BYTES = 4194304          # complex64 [16,64,64,8] cube per frame-sensor
WORKLOADS = {
    "e2e": "end-to-end preproc+inference stream, batch 32 (BASELINE.json configs[2]): int16 ADC of 39 frames x 2 sensors -> "
           "FFT cascade -> window/standardise -> MSCSA-PRGCN forward -> argmax keypoints [32,14,2]",
    "forward-b1": "MSCSA-PRGCN forward-only inference, batch 1 (BASELINE.json configs[1]), mscsa_prgcn.yaml, CUDA-graph replay",
    "cascade": "fft-cascade sweep (BASELINE.json configs[4]): int16 DCA1000 words -> complex64 [16,64,64,8] cubes",
    "cascade-sweep": "fft-cascade throughput sweep 1k-1M radar frames (BASELINE.json configs[4]), frames sharded over the ranks, "
                     "resident chunks of 1024 frames (2048 frame-sensors: 1.5 GiB int16 in, 8 GiB complex64 out)",
    "train": "MSCSA-PRGCN training step (BASELINE.json configs[3] per-GPU shard): train-mode forward + backward + gradient all-reduce + Adam, "
             "synthetic VRDAE inputs and joints",
}
MODEL_FWD_FLOPS = 137.09e9      # per sample, SURVEY.md §8 d (2 x MACs of the reference forward)


def make_cfg():
    ns = types.SimpleNamespace
    return ns(DATASET=ns(numFrames=8, rangeSize=64, heatmapSize=64, azimuthSize=64, elevationSize=8, numGroupFrames=8,
                         numKeypoints=14, imgSize=256),
              MODEL=ns(numFilters=32), TRAINING=ns(lossDecay=-1))


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fp:
            p = json.load(fp)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "25"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            # nvidia-smi can take a second to come up on a fresh box: wait for its first line, then only count the samples taken after
            # mark() — the ones inside the timed regions
            deadline = time.time() + 4.0
            while time.time() < deadline and self._lines() == 0:
                time.sleep(0.02)
        except Exception:
            self.proc = None
        self.skip = self._lines()

    def _lines(self):
        try:
            with open(self.path) as fp:
                return sum(1 for _ in fp)
        except Exception:
            return 0

    def mark(self):
        """Call right before the timed region: samples taken so far (idle GPU) are not reported."""
        self.skip = self._lines()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        try:
            for i, line in enumerate(open(self.path)):
                if i < getattr(self, "skip", 0):
                    continue
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); smax.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            busy = [c for c in sm if c > 0.5 * max(sm)] or sm
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------ CPU reference arm
def _cpu_cascade_worker(seed):
    from oracle import cascade
    frame = cascade.synth_frame(seed, seed & 1)
    t0 = time.perf_counter()
    cascade.generate_heatmap_looped(frame)
    return time.perf_counter() - t0


def cpu_cascade_sample(n_fs, procs):
    """Oracle port with the reference's per-cell np.fft call pattern (process_iwr1843.py:144-164), `procs`-process pool."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        per = pool.map(_cpu_cascade_worker, list(range(n_fs)))
    wall = time.perf_counter() - t0
    return {"wall_s": wall, "per_call_s": statistics.median(per), "fs_per_s": n_fs / wall}


def _cpu_loader_worker(seed):
    import numpy as np
    from oracle import loader
    rng = np.random.default_rng(seed)
    cube = rng.standard_normal((16, 64, 64, 8)) + 1j * rng.standard_normal((16, 64, 64, 8))
    t0 = time.perf_counter()
    loader.vrdae_from_cubes([cube] * 8)
    return time.perf_counter() - t0


def cpu_e2e_sample(cores, forward_reps=2):
    """Reference CPU pipeline cost per radar frame, every stage using all `cores` host threads:
    2 x generateHeatmap (process pool) + loader window/normalise for 2 sensors (process pool) + HuPRNet forward B=1 (torch-CPU)."""
    import multiprocessing as mp
    import torch
    from oracle import model as om
    n_fs = 2 * max(1, cores // 2) if cores > 1 else 2
    casc = cpu_cascade_sample(n_fs, cores)
    ctx = mp.get_context("fork")
    n_win = max(2, cores)
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_loader_worker, list(range(n_win)))
    loader_s_per_sensor_window = (time.perf_counter() - t0) / n_win
    torch.set_num_threads(cores)
    sd = om.make_state_dict(0)
    hori, vert = om.make_vrdae(1, 0)
    with torch.no_grad():
        om.huprnet_forward(sd, hori, vert)
        t0 = time.perf_counter()
        for _ in range(forward_reps):
            om.huprnet_forward(sd, hori, vert)
        fwd_s = (time.perf_counter() - t0) / forward_reps
    per_frame_s = 2.0 / casc["fs_per_s"] + 2.0 * loader_s_per_sensor_window + fwd_s
    sample = ("per radar frame, all %d host threads per stage: 2 cascades (%d frame-sensors through the numpy port with the reference's "
              "per-cell np.fft pattern, %d-process pool: %.3f s/frame) + loader window+Normalize for 2 sensors (%d windows, pool: %.3f s/frame) "
              "+ HuPRNet forward B=1 torch-CPU fp32 (%d reps: %.3f s)" %
              (cores, n_fs, cores, 2.0 / casc["fs_per_s"], n_win, 2.0 * loader_s_per_sensor_window, forward_reps, fwd_s))
    return {"frames_per_s": 1.0 / per_frame_s, "per_frame_s": per_frame_s, "sample": sample,
            "stage_s": {"cascade": 2.0 / casc["fs_per_s"], "loader": 2.0 * loader_s_per_sensor_window, "forward": fwd_s}}


def cpu_baseline_for(workload, cores):
    if workload == "cascade":
        n_fs = 2 * max(1, cores // 2) if cores > 1 else 2
        r = cpu_cascade_sample(n_fs, cores)
        return r["fs_per_s"] / 2.0, ("%d frame-sensors, oracle.cascade.generate_heatmap_looped (numpy port, reference call pattern), "
                                     "%d-process pool, %.1f s wall, %.2f s per call" % (n_fs, cores, r["wall_s"], r["per_call_s"]))
    if workload == "train":
        import numpy as np
        import torch
        from oracle import model as om
        torch.set_num_threads(cores)
        sd = om.make_state_dict(0)
        hori, vert = om.make_vrdae(2, 0)
        joints = np.random.default_rng(0).integers(0, 256, (2, 14, 2))
        t0 = time.perf_counter()
        om.training_gradients(sd, hori, vert, joints)
        dt = time.perf_counter() - t0
        return 2.0 / dt, ("forward + backward of the torch-CPU fp32 oracle (autograd), batch 2, %d threads: %.2f s (optimizer step not "
                          "included)" % (cores, dt))
    r = cpu_e2e_sample(cores)
    if workload == "forward-b1":
        return 1.0 / r["stage_s"]["forward"], "HuPRNet forward B=1, oracle.model (torch-CPU fp32 restatement), %d threads" % cores
    return r["frames_per_s"], r["sample"]


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals = []
    t_all = time.perf_counter()
    for i in range(args.warmup_ref + args.steps_ref):
        t0 = time.perf_counter()
        fps, sample = cpu_baseline_for(args.workload, cores)
        if i >= args.warmup_ref:
            vals.append((fps, time.perf_counter() - t0, sample))
    fps = statistics.median([v[0] for v in vals])
    line = {"metric": METRIC, "value": fps, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps_ref, "warmup": args.warmup_ref, "ms_per_step": statistics.median([v[1] for v in vals]) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (cascade c128)", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload]},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": vals[-1][2]},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def summarise_profile(rec):
    fam = {}
    for name, flops, ms in rec:
        # the operand-plane pass of the two-unit convolutions (ops.quantize_planes) is part of the convolution family's time
        key = "conv_gemm" if name.startswith("conv_gemm") or name == "quantize_planes" else name
        f = fam.setdefault(key, {"ms": 0.0, "flops": 0.0, "launch_calls": 0, "unit_flops": 0.0})
        f["ms"] += ms; f["flops"] += flops; f["launch_calls"] += 1
        f["unit_flops"] += flops * (2 if name.endswith(".q") else 3)      # tensor units per k-step: 2 (fp16 + 2 e4m3 at twice the rate) or 3 bf16
    sub = {}
    for name, flops, ms in rec:
        if name.startswith("conv_gemm"):
            s = sub.setdefault(name, {"ms": 0.0, "flops": 0.0, "calls": 0})
            s["ms"] += ms; s["flops"] += flops; s["calls"] += 1
    return fam, sub


def count_graph_kernels(replay):
    """Kernel launches of ONE graph replay by origin, from a CUPTI trace (torch.profiler): {"hupr": n, "other": m, "other_names": [...]}.
    ``other`` are kernels that do not come from libhupr_b200.so (framework element-wise kernels, NCCL); memsets / memcpys are DMA nodes,
    counted separately.  None if no trace can be taken on this box."""
    import torch
    try:
        from torch.profiler import ProfilerActivity, profile
        replay()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            replay()
            torch.cuda.synchronize()
        hupr, other, memops, names = 0, 0, 0, {}
        for ev in prof.events():
            if ev.device_type != torch.autograd.DeviceType.CUDA:
                continue
            name = ev.name
            if name.startswith("Memset") or name.startswith("Memcpy"):
                memops += 1
            elif "hupr::" in name:
                hupr += 1
            else:
                other += 1
                names[name[:80]] = names.get(name[:80], 0) + 1
        if hupr + other == 0:
            return None
        return {"hupr": hupr, "other": other, "memset_memcpy_nodes": memops,
                "other_names": sorted(names.items(), key=lambda kv: -kv[1])[:8]}
    except Exception as exc:      # no CUPTI on the box, or the profiler refuses graphs
        return {"error": str(exc)[:200]}


def train_probe(dev, world, gen, steps, products):
    """Data-parallel training evidence inside the default multi-rank run (BASELINE.json configs[3]; reference semantics
    /root/reference/tools/run.py:74-79): graph-replayed forward+backward at 32 samples per GPU, ONE NCCL all-reduce over the flat
    gradient buffer, Adam.  Returns per-step and exposed-collective times (CUDA events, max over ranks)."""
    import torch
    import torch.distributed as dist
    from hupr_b200.models import HuPRNet
    from hupr_b200.training import TrainStep
    torch.manual_seed(0)
    model = HuPRNet(make_cfg()).to(dev).train()
    trainer = TrainStep(model, products=products)
    b = 32
    hori = torch.randn((b, 8, 8, 2, 64, 64, 8), generator=gen, device=dev)
    vert = torch.randn((b, 8, 8, 2, 64, 64, 8), generator=gen, device=dev)
    joints = torch.randint(0, 256, (b, 14, 2), generator=gen, device=dev)
    replay = trainer.capture(hori, vert, joints, with_optimizer=False)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ar_ms, losses = [], []

    def one(timed):
        loss, _ = replay()
        if timed:
            ev[1].record()
        trainer.all_reduce_gradients()
        if timed:
            ev[2].record()
        trainer.optimizer_step()
        return loss
    for _ in range(2):
        one(False)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(steps):
        loss = one(True)
        torch.cuda.current_stream().synchronize()
        losses.append(float(loss))                # the graph's loss tensor is static: read it before the next replay overwrites it
        ar_ms.append(ev[1].elapsed_time(ev[2]))
    ev[3].record()
    torch.cuda.synchronize()
    t = torch.tensor([ev[0].elapsed_time(ev[3]) / steps, sum(ar_ms) / len(ar_ms)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # replicas must stay identical: the all-reduced gradient (and so the weights) agree across ranks to the bit
    chk = torch.stack([trainer.flat_p.double().sum(), trainer.flat_p.double().abs().max()])
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ms, ar = float(t[0]), float(t[1])
    return {"what": "graph-replayed train step (forward+backward) + ONE NCCL sum all-reduce of the flat fp32 gradient buffer + Adam",
            "samples_per_gpu": b, "global_batch": b * world, "steps": steps, "ms_per_step": ms, "allreduce_ms": ar,
            "allreduce_bytes": trainer.flat_g.numel() * 4, "samples_per_s": b * world / (ms * 1e-3),
            "products_per_k_step": products, "replicas_identical": bool(torch.equal(lo, hi)),
            "loss_first_last": [losses[0], losses[-1]]}


def run_cascade_sweep(args, dev, world, rank, local_rank, peaks, gen, sync_all):
    """BASELINE.json configs[4]: N radar frames (hori + vert) through hupr_fft_cascade_i16, N/world per rank, in resident chunks.
    The chunk's int16 words are generated once on the device and re-used for every chunk (the arithmetic does not depend on the
    data; 1 M frames would be 1.5 TB of distinct input); every chunk is a real launch reading 1.5 GiB and writing 8 GiB."""
    import torch
    import torch.distributed as dist
    from hupr_b200.preprocessing.process_iwr1843 import cascade_i16, FRAME_WORDS
    chunk = 1024                                              # radar frames per launch
    adc = torch.randint(-2048, 2048, (2 * chunk, FRAME_WORDS), generator=gen, dtype=torch.int16, device=dev)
    cube = torch.empty((2 * chunk, 16, 64, 64, 8), dtype=torch.complex64, device=dev)
    for _ in range(max(args.warmup, 3)):
        cascade_i16(adc, cube)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    table = []
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for total in (1024, 4096, 16384, 65536, 262144, 1048576):
        mine = total // world                                 # frames of this rank (contiguous range; no collective on the data path)
        full, tail = divmod(mine, chunk)
        sync_all()
        start.record()
        for _ in range(full):
            cascade_i16(adc, cube)
        if tail:
            cascade_i16(adc[:2 * tail], cube[:2 * tail])
        stop.record()
        sync_all()
        t = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        table.append({"frames": total, "frames_per_gpu": mine, "launches_per_gpu": full + (1 if tail else 0), "ms": ms,
                      "frames_per_s": total / (ms * 1e-3),
                      "gbs_per_gpu": 2 * mine * (FS_IN_BYTES + FS_OUT_BYTES) / (ms * 1e-3) / 1e9})
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        last = table[-1]
        line = {"metric": METRIC, "value": last["frames_per_s"], "unit": UNIT, "n_gpus": world, "steps": 1, "warmup": max(args.warmup, 3),
                "ms_per_step": last["ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic (seeded random int16 ADC words, one resident chunk re-used)",
                "config": {"workload": WORKLOADS["cascade-sweep"], "chunk_frames": chunk,
                           "l2": "each launch streams 9.5 GiB, far above the 126 MB L2", "parallelism": "frames sharded across ranks, no collective"},
                "roofline": {"bound": "hbm", "achieved": last["gbs_per_gpu"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": last["gbs_per_gpu"] / peaks["hbm_gbs"], "traffic": None, "kernel": "hupr::cascade_kernel",
                             "peak_source": peaks["source"]},
                "sweep": table, "cpu_baseline": None, "gpu_launches": sum(r["launches_per_gpu"] for r in table), "clocks": clocks,
                "e2e": None}
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="e2e", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="e2e: poses (windows) per GPU per step (default 32); train: samples per GPU (default 32 = the per-GPU shard of configs[3])")
    ap.add_argument("--frames-per-step", type=int, default=1024, help="cascade: radar frames per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ingest", default="adc", choices=["adc", "vrdae"],
                    help="train e2e leg: host buffers hold raw int16 ADC words (cascade + standardisation on the device, default) or float32 VRDAE maps")
    ap.add_argument("--no-train-probe", action="store_true", help="multi-rank e2e runs: skip the data-parallel training probe record")
    ap.add_argument("--single-bf16", action="store_true", help="one bf16 product per k-step instead of the fp32-equivalent 3-product split")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 32
    args.steps_ref = max(1, min(args.steps, 2))
    args.warmup_ref = 0

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from hupr_b200 import ops
    from hupr_b200.models import HuPRNet
    from hupr_b200.pipeline import RadarPoseStream
    from hupr_b200.preprocessing.process_iwr1843 import cascade_i16, FRAME_WORDS

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (B200); there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    if args.workload == "cascade-sweep":
        run_cascade_sweep(args, dev, world, rank, local_rank, peaks, gen, sync_all)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- build the workload ---------------------------------------------------------------------------------------------
    profile_step = None
    if args.workload == "cascade":
        units = args.frames_per_step
        n_fs = 2 * units
        adc = torch.randint(-2048, 2048, (n_fs, FRAME_WORDS), generator=gen, dtype=torch.int16, device=dev)
        cube = torch.empty((n_fs, 16, 64, 64, 8), dtype=torch.complex64, device=dev)
        step = lambda: cascade_i16(adc, cube)
        launches_per_step = 1
        n_e2e = 128
        h_in = torch.randint(-2048, 2048, (2 * n_e2e, FRAME_WORDS), dtype=torch.int16).pin_memory()
        h_out = torch.empty((2 * n_e2e, 16, 64, 64, 8), dtype=torch.complex64).pin_memory()
        d_in = torch.empty_like(h_in, device=dev)
        d_out = torch.empty((2 * n_e2e, 16, 64, 64, 8), dtype=torch.complex64, device=dev)

        def e2e_step():
            d_in.copy_(h_in, non_blocking=True)
            cascade_i16(d_in, d_out)
            h_out.copy_(d_out, non_blocking=True)
        e2e_units, h2d, d2h = n_e2e, 2 * n_e2e * FS_IN_BYTES, 2 * n_e2e * FS_OUT_BYTES
        l2_note = "inputs+outputs (%.1f GB) exceed the 126 MB L2" % (n_fs * (FS_IN_BYTES + FS_OUT_BYTES) / 1e9)
        dtype = "f32"
    elif args.workload == "train":
        from hupr_b200.training import TrainStep
        torch.manual_seed(0)
        model = HuPRNet(make_cfg()).to(dev).train()
        products = 1 if args.single_bf16 else 3
        trainer = TrainStep(model, products=products)
        units = args.batch
        hori = torch.randn((units, 8, 8, 2, 64, 64, 8), generator=gen, device=dev)
        vert = torch.randn((units, 8, 8, 2, 64, 64, 8), generator=gen, device=dev)
        joints = torch.randint(0, 256, (units, 14, 2), generator=gen, device=dev)
        h_h, h_v, h_j = hori.cpu().pin_memory(), vert.cpu().pin_memory(), joints.cpu().pin_memory()
        h_loss = torch.empty(2, dtype=torch.float32).pin_memory()

        before = ops.launch_count()
        trainer.forward_backward(hori, vert, joints)
        trainer.optimizer_step()
        launches_per_step = ops.launch_count() - before
        # the step is ~1 150 launches: replay it as a CUDA graph (single rank: including Adam; several ranks: the NCCL all-reduce and
        # the Adam launch run eagerly after the captured forward+backward)
        replay = trainer.capture(hori, vert, joints, with_optimizer=(world == 1))

        def step():
            out = replay()
            if world > 1:
                trainer.all_reduce_gradients()
                trainer.optimizer_step()
            return out

        if args.ingest == "adc":
            # compact ingest: raw int16 DCA1000 words of each sample's 8-frame window (12.6 MB per sample); FFT cascade + standardisation on the device
            h_ah = torch.randint(-2048, 2048, (units, 8, FRAME_WORDS), dtype=torch.int16).pin_memory()
            h_av = torch.randint(-2048, 2048, (units, 8, FRAME_WORDS), dtype=torch.int16).pin_memory()
            upload = lambda: trainer.prefetch_adc(h_ah, h_av, h_j)
            take = lambda: trainer.take_prefetched_adc(hori, vert, joints)
            h2d = 2 * h_ah.numel() * 2 + h_j.numel() * 8
        else:
            upload = lambda: trainer.prefetch(h_h, h_v, h_j)
            take = lambda: trainer.take_prefetched(hori, vert, joints)
            h2d = 2 * h_h.numel() * 4 + h_j.numel() * 8
        upload()

        def e2e_step():
            # public ingest API: the upload of the NEXT batch (copy stream, pinned host tensors) overlaps this step's compute; every step
            # still moves its own inputs host->device and its two loss scalars device->host inside the timed region
            take()
            upload()
            l, l2 = step()
            h_loss[0:1].copy_(l.reshape(1), non_blocking=True)
            h_loss[1:2].copy_(l2.reshape(1), non_blocking=True)
        e2e_units, d2h = units, 8
        profile_step = lambda: trainer.forward_backward(hori, vert, joints)
        l2_note = "per-step working set (saved activations + gradients, tens of GB) exceeds the 126 MB L2"
        dtype = ("bf16 tensor-core products, fp32 accumulate, in every convolution / data-gradient / weight-gradient launch (attention and "
                 "element-wise kernels keep hi/lo inputs), fp32 master weights / Adam" if args.single_bf16 else
                 "bf16 tensor-core products, fp32 accumulate (3-product hi/lo split: fp32-equivalent), fp32 master weights / Adam")
    else:
        torch.manual_seed(0)
        model = HuPRNet(make_cfg(), split=not args.single_bf16).to(dev).eval()      # random-init weights of the reference architecture
        dtype = "bf16 tensor-core products, fp32 accumulate" + ("" if args.single_bf16 else " (3-product hi/lo split: fp32-equivalent)")
        if not args.single_bf16 and model.quant_cross_terms:
            dtype += ("; the 3-tap convolutions with cout % 128 == 0 use 2 tensor units per k-step instead (fp16 main product + the two hi*lo cross "
                      "terms as e4m3 products, fp32 accumulate; HUPR_QUANT=0 turns this off) — heat maps within the same 1e-3 bar, keypoints identical")
        if args.workload == "e2e":
            units = args.batch
            stream = RadarPoseStream(model, units, dev).prepare()
            stream.adc.copy_(torch.randint(-2048, 2048, stream.adc.shape, generator=gen, dtype=torch.int16, device=dev))
            step = stream.step
            launches_per_step = stream.launches_per_step
            h_adc = torch.randint(-2048, 2048, (2, stream.n_frames, FRAME_WORDS), dtype=torch.int16).pin_memory()
            h_kp = torch.empty((units, 14, 2), dtype=torch.float32).pin_memory()

            stream.prefetch(h_adc[0], h_adc[1])

            def e2e_step():
                # public streaming API: the upload of the NEXT 39x2 frames (copy stream) overlaps this step's compute; every step still
                # moves its own 61 MB host->device and its keypoints device->host inside the timed region
                kp = stream.step_prefetched()
                stream.prefetch(h_adc[0], h_adc[1])
                h_kp.copy_(kp, non_blocking=True)
            e2e_units, h2d, d2h = units, h_adc.numel() * 2, h_kp.numel() * 4
            profile_step = stream._step_eager
            l2_note = "per-step working set (activations, several GB) exceeds the 126 MB L2"
        else:   # forward-b1
            units = 1
            hori = torch.randn((1, 8, 8, 2, 64, 64, 8), generator=gen, device=dev)
            vert = torch.randn((1, 8, 8, 2, 64, 64, 8), generator=gen, device=dev)
            kp = torch.empty((1, 14, 2), dtype=torch.float32, device=dev)

            def eager():
                heat, gcn = model(hori, vert)
                ops.keypoints_argmax(gcn.view(1, 14, 64, 64), kp)
            eager()
            before = ops.launch_count()
            eager()
            launches_per_step = ops.launch_count() - before
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                eager()
            flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

            def step():
                flush.zero_()          # evict L2 between timed iterations (weights + activations would otherwise stay resident)
                graph.replay()
            h_h, h_v = hori.cpu().pin_memory(), vert.cpu().pin_memory()
            h_kp = torch.empty((1, 14, 2), dtype=torch.float32).pin_memory()

            def e2e_step():
                hori.copy_(h_h, non_blocking=True)
                vert.copy_(h_v, non_blocking=True)
                graph.replay()
                h_kp.copy_(kp, non_blocking=True)
            e2e_units, h2d, d2h = 1, 2 * h_h.numel() * 4, h_kp.numel() * 4
            profile_step = eager
            l2_note = "L2 flushed (256 MiB memset) between timed iterations; the flush is inside the timed region"

    # ---- timed region: value (inputs resident in HBM) -------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    sync_all()
    if rank == 0:
        sampler.mark()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        step()
    stop.record()
    sync_all()
    t = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = world * units / (ms_per_step * 1e-3)

    # ---- e2e leg: host (pinned) inputs, H2D + step + D2H of the result inside the timed region --------------------------
    for _ in range(2):
        e2e_step()
    sync_all()
    e_steps = max(3, min(args.steps, 10))
    start.record()
    for _ in range(e_steps):
        e2e_step()
    stop.record()
    sync_all()
    t = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_units / (float(t.item()) / e_steps * 1e-3)
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel-family breakdown (one instrumented eager step, CUDA events around every library call) ---------------
    breakdown = None
    if args.workload == "cascade":
        launch_bytes = 2 * units * (FS_IN_BYTES + FS_OUT_BYTES)
        achieved = launch_bytes / (ms_per_step * 1e-3) / 1e9
        # DRAM traffic per frame-sensor from the committed ncu --set full capture (profiles/r01_cascade_metrics.csv: 465.6 MB read +
        # 2 423.5 MB written for a 592-frame-sensor launch = 4 880 262 B each, vs 4 980 736 B algorithmic), scaled to this launch
        traffic = 2 * units * (465608960 + 2423526000) / 592.0
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                    "traffic": traffic, "kernel": "hupr::cascade_kernel", "peak_source": peaks["source"],
                    "algorithmic_bytes_per_launch": launch_bytes}
    else:
        profile_step()
        ops.profile_begin()
        profile_step()
        fam, sub = summarise_profile(ops.profile_end())
        total_ms = sum(f["ms"] for f in fam.values())
        conv = fam["conv_gemm"]
        achieved = conv["flops"] / (conv["ms"] * 1e-3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        # DRAM traffic of the family's dominant kernel from its ncu --set full capture, ONE launch of the 64 -> 128 channel 3x3x3 layer at
        # batch 32.  Default arithmetic (profiles/r02_halo128_q64_pairs_metrics.csv, conv_halo_kernel<128,2,true>: CTA pairs, two-unit
        # products, output stored as hi/lo planes AND as the next layer's fp16 + e4m3 operand planes): 270.5 MB read + 1 019.9 MB written
        # against 269.3 MB of input + weights and 1 073.7 MB of output.  Three bf16 products (profiles/r02_halo128_pairs_metrics.csv,
        # conv_halo_kernel<128,3,true>): 270.3 MB + 487.7 MB against 269.3 + 536.9 MB.  No DRAM re-reads either way: the 27-tap operand
        # re-reads are served by L2 (part of the output is still in L2 at exit).
        two_unit = (not args.single_bf16) and any(k.endswith(".q") for k in sub)
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": (270486272.0 + 1019883000.0) if two_unit else (270325248.0 + 487689728.0),
                    "traffic_algorithmic_bytes": (269317120.0 + 1073741824.0) if two_unit else (269317120.0 + 536870912.0),
                    "traffic_of": "one conv_halo_kernel<128,%s,true> launch (64->128 channels, 3x3x3, batch 32), ncu dram__bytes_read.sum + "
                                  "dram__bytes_write.sum" % ("2" if two_unit else "3"),
                    "kernel": "hupr::conv_gemm_kernel (all %d launches of one step; FLOPs = 2 x MACs of the fp32 contraction — the "
                              "tensor pipe executes %s that)" % (conv["launch_calls"], "1x" if args.single_bf16 else
                                                                "3x (bf16 hi/lo products), 2x on the two-unit 3-tap convolutions (fp16 + e4m3 cross terms)"),
                    "peak_source": peaks["source"] + " (sustained cuBLAS bf16)", "algorithmic_flops_per_step": conv["flops"],
                    "kernel_ms_per_step": conv["ms"], "share_of_step": conv["ms"] / total_ms,
                    "executed_tensor_tflops": achieved if args.single_bf16 else conv["unit_flops"] / (conv["ms"] * 1e-3) / 1e12,
                    "two_unit_share_of_flops": sum(v["flops"] for k, v in sub.items() if k.endswith(".q")) / max(conv["flops"], 1.0)}
        # the other named kernel families of this step against their own rooflines (algorithmic work / CUDA-event time of the family)
        families = {}
        for key in ("attention_fwd", "attention_bwd", "conv_wgrad", "matmul_tn"):
            if key in fam and fam[key]["flops"] > 0:
                tf = fam[key]["flops"] / (fam[key]["ms"] * 1e-3) / 1e12
                three = key.startswith("attention") or not args.single_bf16
                families[key] = {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                                 "executed_frac": tf * (3 if three else 1) / peak, "ms": fam[key]["ms"]}
        if "fft_cascade_i16" in fam and args.workload == "e2e":
            n_fs = 2 * (units + 7)
            gbs = n_fs * (FS_IN_BYTES + FS_OUT_BYTES) / (fam["fft_cascade_i16"]["ms"] * 1e-3) / 1e9
            families["fft_cascade_i16"] = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                                           "ms": fam["fft_cascade_i16"]["ms"], "frame_sensors": n_fs}
        roofline["families"] = families
        breakdown = {k: {"ms": round(v["ms"], 4), "share": round(v["ms"] / total_ms, 4), "calls": v["launch_calls"]}
                     for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
        breakdown["conv_gemm_by_shape"] = {k: {"ms": round(v["ms"], 4), "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2), "calls": v["calls"]}
                                           for k, v in sorted(sub.items(), key=lambda kv: -kv[1]["ms"])}

    kernels = None
    if args.workload == "train" and rank == 0:
        kernels = count_graph_kernels(replay)
    probe = None
    if world > 1 and args.workload == "e2e" and not args.no_train_probe:
        probe = train_probe(dev, world, gen, 5, 1 if args.single_bf16 else 3)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            fps, sample = cpu_baseline_for(args.workload, cores)
            cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic (seeded random int16 ADC words; random-init weights of the reference architecture)",
            "config": {"workload": WORKLOADS[args.workload] + ((" [e2e ingest: %s]" % args.ingest) if args.workload == "train" else ""),
                       "units_per_gpu_per_step": units, "l2": l2_note,
                       "parallelism": ("data parallel: samples sharded across ranks, one NCCL sum all-reduce over the flat fp32 gradient buffer per step"
                                       if args.workload == "train" else "frames sharded across ranks, no collective")},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "units_per_gpu_per_step": e2e_units},
            "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
            "clocks": clocks,
        }
        if breakdown is not None:
            line["breakdown"] = breakdown
        if kernels is not None:
            line["kernels_per_replay"] = kernels
        if probe is not None:
            line["train_probe"] = probe
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
