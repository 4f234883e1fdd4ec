"""CPU multi-process test (gloo, world_size 2 and 3) of the N>1 host logic: frame/window sharding and the keypoint gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_windows, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hupr_b200 import sharding
    first, count = sharding.window_shard(n_windows, world, rank)
    f0, nf = sharding.frames_for_windows(first, count)
    assert f0 == first and nf == (count + 7 if count else 0)
    # stand-in for the per-rank pipeline output: keypoint k of window b is (b, k) -> any mis-ordering shows in the gather
    b = torch.arange(first, first + count, dtype=torch.float32).view(-1, 1, 1).expand(count, 14, 1)
    k = torch.arange(14, dtype=torch.float32).view(1, 14, 1).expand(count, 14, 1)
    full = sharding.gather_keypoints(torch.cat([b, k], dim=2).contiguous(), n_windows)
    torch.save(full, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_windows", [(2, 64), (2, 33), (3, 10)])
def test_window_shards_cover_stream_and_gather_in_order(world, n_windows, tmp_path):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    mp.spawn(_worker, args=(world, _free_port(), n_windows, str(tmp_path)), nprocs=world, join=True)
    expect_b = torch.arange(n_windows, dtype=torch.float32)
    for r in range(world):
        full = torch.load(os.path.join(str(tmp_path), "rank%d.pt" % r))
        assert full.shape == (n_windows, 14, 2)
        assert torch.equal(full[:, 0, 0], expect_b) and torch.equal(full[0, :, 1], torch.arange(14, dtype=torch.float32))


def test_window_shard_partition_properties():
    from hupr_b200 import sharding
    for n in (0, 1, 7, 32, 1000):
        for world in (1, 2, 4, 8):
            spans = [sharding.window_shard(n, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == n
            assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        sharding.window_shard(4, 2, 2)


def _grad_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hupr_b200 import sharding
    torch.manual_seed(0)
    w = torch.randn(5, 3, dtype=torch.float64, requires_grad=True)          # identical replicas
    x = torch.randn(8, 5, dtype=torch.float64)                              # the global batch, same on every rank
    y = torch.randn(8, 3, dtype=torch.float64)
    per = 8 // world
    xs, ys = x[rank * per:(rank + 1) * per], y[rank * per:(rank + 1) * per]
    loss = torch.nn.functional.binary_cross_entropy(torch.sigmoid(xs @ w), torch.sigmoid(ys))      # mean over the LOCAL shard
    loss.backward()
    flat = w.grad.reshape(-1).clone()
    sharding.average_gradients(flat)
    torch.save(flat, os.path.join(out_dir, "grad%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_averaged_shard_gradients_equal_global_batch_gradient(world, tmp_path):
    """The training collective (TrainStep.all_reduce_gradients -> sharding.average_gradients): per-rank mean losses over equal shards,
    one sum all-reduce, divide by the world size == the gradient of the mean loss over the whole batch on one device."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    mp.spawn(_grad_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    torch.manual_seed(0)
    w = torch.randn(5, 3, dtype=torch.float64, requires_grad=True)
    x = torch.randn(8, 5, dtype=torch.float64)
    y = torch.randn(8, 3, dtype=torch.float64)
    torch.nn.functional.binary_cross_entropy(torch.sigmoid(x @ w), torch.sigmoid(y)).backward()
    for r in range(world):
        flat = torch.load(os.path.join(str(tmp_path), "grad%d.pt" % r))
        assert float((flat - w.grad.reshape(-1)).abs().max()) < 1e-14


def test_average_gradients_is_a_no_op_without_a_process_group():
    from hupr_b200 import sharding
    flat = torch.arange(4, dtype=torch.float32)
    assert torch.equal(sharding.average_gradients(flat.clone()), flat)
