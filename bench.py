#!/usr/bin/env python
"""bench.py — headline benchmark of the hupr_b200 hot path (contract: see the task statement / DESIGN.md §bench).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cascade|e2e] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): radar frames/s (one radar frame = one hori + one vert IWR1843 capture).
A "step" is one pass of the hot path over one batch of synthetic input resident in HBM.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "radar_frames_per_sec"
UNIT = "frames/s"
FS_IN_BYTES = 786432            # int16 ADC words per frame-sensor
FS_OUT_BYTES = 4194304          # complex64 [16,64,64,8] cube per frame-sensor


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fp:
            p = json.load(fp)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback"}


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
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

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
            for line in open(self.path):
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
def _cpu_worker(seed):
    from oracle import cascade
    frame = cascade.synth_frame(seed, seed & 1)
    t0 = time.perf_counter()
    cascade.generate_heatmap_looped(frame)
    return time.perf_counter() - t0


def cpu_cascade_sample(n_fs, procs):
    """Time the oracle's reference-shaped port (same per-cell np.fft call pattern as
    process_iwr1843.py:144-164) on `n_fs` frame-sensors with a `procs`-process pool.  Returns frames/s."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        per = pool.map(_cpu_worker, list(range(n_fs)))
    wall = time.perf_counter() - t0
    return {"frames_per_s": (n_fs / 2.0) / wall, "wall_s": wall, "per_call_s": statistics.median(per)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_fs = 2 * max(1, cores // 2) if cores > 1 else 2
    vals = []
    for i in range(args.warmup_ref + args.steps_ref):
        r = cpu_cascade_sample(n_fs, cores)
        if i >= args.warmup_ref:
            vals.append(r)
    fps = statistics.median([v["frames_per_s"] for v in vals])
    wall = statistics.median([v["wall_s"] for v in vals])
    sample = ("%d frame-sensors per step through oracle.cascade.generate_heatmap_looped (numpy port with the reference's "
              "per-cell np.fft call pattern), %d-process pool; median per-call %.2f s" % (n_fs, cores, vals[-1]["per_call_s"]))
    line = {"metric": METRIC, "value": fps, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps_ref, "warmup": args.warmup_ref, "ms_per_step": wall * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": {"workload": args.workload_name, "frame_sensors_per_step": n_fs},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cascade", choices=["cascade"])
    ap.add_argument("--frames-per-step", type=int, default=1024, help="radar frames (hori+vert pairs) per GPU per step")
    ap.add_argument("--e2e-frames", type=int, default=128, help="radar frames per GPU per step in the host-buffer e2e leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.workload_name = "fft-cascade sweep (BASELINE.json configs[4]): int16 DCA1000 words -> complex64 [16,64,64,8] cubes"
    args.steps_ref = max(1, min(args.steps, 3))
    args.warmup_ref = 1 if args.warmup > 0 else 0

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
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

    n_fs = 2 * args.frames_per_step                       # frame-sensors per GPU per step (hori + vert)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    adc = torch.randint(-2048, 2048, (n_fs, FRAME_WORDS), generator=gen, dtype=torch.int16, device=dev)
    cube = torch.empty((n_fs, 16, 64, 64, 8), dtype=torch.complex64, device=dev)

    def step():
        cascade_i16(adc, cube)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    sync_all()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        step()
    stop.record()
    sync_all()
    elapsed_ms = start.elapsed_time(stop)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    ms_per_step = elapsed_ms / args.steps
    value = world * args.frames_per_step / (ms_per_step * 1e-3)
    launch_bytes = n_fs * (FS_IN_BYTES + FS_OUT_BYTES)
    achieved = launch_bytes / (ms_per_step * 1e-3) / 1e9     # one cascade launch per step

    # ---- e2e leg: host (pinned) int16 in, H2D + kernel + D2H of the result inside the timed region ----
    n_e2e = 2 * args.e2e_frames
    h_in = torch.randint(-2048, 2048, (n_e2e, FRAME_WORDS), dtype=torch.int16).pin_memory()
    h_out = torch.empty((n_e2e, 16, 64, 64, 8), dtype=torch.complex64).pin_memory()
    d_in = torch.empty_like(h_in, device=dev)
    d_out = torch.empty((n_e2e, 16, 64, 64, 8), dtype=torch.complex64, device=dev)

    def e2e_step():
        d_in.copy_(h_in, non_blocking=True)
        cascade_i16(d_in, d_out)
        h_out.copy_(d_out, non_blocking=True)

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
    e2e_value = world * args.e2e_frames / (float(t.item()) / e_steps * 1e-3)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_cpu = 2 * max(1, cores // 2) if cores > 1 else 2
            r = cpu_cascade_sample(n_cpu, cores)
            cpu = {"value": r["frames_per_s"], "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "%d frame-sensors, oracle.cascade.generate_heatmap_looped (numpy port, reference call pattern), "
                             "%d-process pool, %.1f s wall, %.2f s per call" % (n_cpu, cores, r["wall_s"], r["per_call_s"])}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload_name, "frames_per_gpu_per_step": args.frames_per_step,
                       "frame_sensors_per_gpu_per_step": n_fs, "l2": "inputs+outputs (%.1f GB) exceed the 126 MB L2" % (launch_bytes / 1e9),
                       "parallelism": "frames sharded across ranks, no collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": None, "kernel": "hupr::cascade_kernel",
                         "peak_source": peaks["source"], "algorithmic_bytes_per_launch": launch_bytes},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n_e2e * FS_IN_BYTES,
                    "d2h_bytes_per_step": n_e2e * FS_OUT_BYTES, "frames_per_gpu_per_step": args.e2e_frames},
            "gpu_launches": args.steps,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
