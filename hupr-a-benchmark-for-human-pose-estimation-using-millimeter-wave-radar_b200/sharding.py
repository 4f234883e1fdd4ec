"""Frame sharding across GPUs (SURVEY.md §8 e): the hot path partitions by radar frame — every frame-sensor of the FFT cascade
and every 8-frame window of the inference stream is independent (process_iwr1843.py:189-195 and tools/run.py:36-58 carry no
state between frames) — so ranks own contiguous window ranges (plus the 7-frame history their first window needs) and there is
NO data-path collective.  The only exchange is the host-side gather of the ``[n,14,2]`` keypoints for the results file
(tools/base.py:124-152), done here with ``torch.distributed.all_gather`` on whatever backend the job runs (NCCL on GPUs, gloo
in the CPU tests)."""
import torch
import torch.distributed as dist

GROUP = 8


def window_shard(n_windows, world, rank):
    """Contiguous balanced split of ``n_windows`` poses: returns (first_window, count) of ``rank``."""
    if world < 1 or not 0 <= rank < world or n_windows < 0:
        raise ValueError("bad shard request: n_windows=%d world=%d rank=%d" % (n_windows, world, rank))
    base, extra = divmod(n_windows, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def frames_for_windows(first_window, count, group=GROUP):
    """Stream frames a rank must hold for its windows: window b covers frames b .. b+group-1 -> (first_frame, n_frames)."""
    if count == 0:
        return first_window, 0
    return first_window, count + group - 1


def gather_keypoints(local, n_windows, world=None, rank=None):
    """All-gather per-rank keypoints ``[count_r, 14, 2]`` into the stream-ordered ``[n_windows, 14, 2]`` tensor (every rank)."""
    if world is None:
        world, rank = dist.get_world_size(), dist.get_rank()
    counts = [window_shard(n_windows, world, r)[1] for r in range(world)]
    if local.shape[0] != counts[rank]:
        raise ValueError("rank %d holds %d windows, expected %d" % (rank, local.shape[0], counts[rank]))
    width = max(counts) if counts else 0
    padded = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def average_gradients(flat):
    """Data-parallel training (SURVEY.md §8 e): ONE sum all-reduce over the flat gradient buffer, then divide by the world size —
    every rank computed the mean loss of its own equally sized shard, so the result is the gradient of the global mean loss
    (tools/run.py:76-79 run on one device with the whole batch).  No-op without an initialised process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(dist.get_world_size())
    return flat
