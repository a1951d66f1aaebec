"""Data-parallel host logic on CPU with the gloo backend, world_size 2: ray sharding covers every ray
exactly once, the bucketed all-reduce equals the single-process sum, and a sharded step whose means use
the GLOBAL counts (ray count, all-reduced masked-mean denominator) reproduces the full-batch gradient by a
plain sum - also when one rank holds every member of the masked mean."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from semnerf_b200 import dist as snb_dist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_range_partitions_every_ray_once():
    for n in (0, 1, 7, 1024, 65536, 640000):
        for world in (1, 2, 3, 4, 8):
            seen = 0
            prev = 0
            sizes = []
            for r in range(world):
                lo, hi = snb_dist.shard_range(n, r, world)
                assert lo == prev and hi >= lo
                prev = hi
                seen += hi - lo
                sizes.append(hi - lo)
            assert seen == n and prev == n and max(sizes) - min(sizes) <= 1


def test_bucket_ranges_cover_flat_buffer_on_tensor_boundaries():
    from semnerf_b200.model import RSSemanticNeRFB200
    from oracle import render_oracle as O
    from tests.helpers import make_cfgs
    m = RSSemanticNeRFB200(make_cfgs(O.ModelSpec(), 64, 0.05), type("D", (), {"semantic_n_classes": 6})())
    n = m.flat.numel()
    ranges = snb_dist.bucket_ranges(m.table, n, 3)
    assert len(ranges) == 3
    flat = sorted(ranges)
    assert flat[0][0] == 0 and flat[-1][1] == n
    for (a, b), (c, d) in zip(flat[:-1], flat[1:]):
        assert b == c
    offsets = {off for _, off, _ in m.table} | {n}
    assert all(lo in offsets and hi in offsets for lo, hi in ranges)
    # backward produces the heads (end of the buffer) first: that bucket is launched first
    assert ranges[0][1] == n
    sizes = [hi - lo for lo, hi in ranges]
    assert max(sizes) < 0.6 * n


def _worker(rank, world, port, n_params, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, l, w = snb_dist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    # a toy "render": per-ray loss = (w . x_ray)^2 ; global batch of 10 rays sharded across ranks
    x = torch.randn(10, n_params)
    wgt = torch.randn(n_params, requires_grad=True)
    lo, hi = snb_dist.shard_range(10, rank, world)
    # the trainer's convention: every rank normalises by the GLOBAL counts (global ray count, all-reduced masked-mean
    # denominators), so the per-rank losses are shares of the global loss and the gradients simply SUM
    mask = torch.arange(10) < 3                                   # a masked mean whose members all sit in rank 0's shard
    cnt_local = mask[lo:hi].sum().float().view(1)
    cnt = snb_dist.all_reduce_scalar_sum(cnt_local.clone())
    per_ray = (x[lo:hi] @ wgt) ** 2
    loss_local = per_ray.sum() / 10 + (per_ray * mask[lo:hi]).sum() / cnt.clamp_min(1)
    loss_local.backward()
    g = wgt.grad.clone()
    red = snb_dist.GradAllReducer([(0, n_params // 3), (n_params // 3, n_params)][::-1])
    red.launch(g)
    red.wait()
    wfull = wgt.detach().clone().requires_grad_(True)
    pr = (x @ wfull) ** 2
    (pr.mean() + pr[mask].mean()).backward()
    ok = torch.allclose(g, wfull.grad, rtol=1e-5, atol=1e-6) and cnt.item() == 3
    cnt = snb_dist.all_reduce_scalar_sum(torch.tensor([float(hi - lo)]))
    ok = ok and cnt.item() == 10
    snb_dist.barrier()
    open(os.path.join(tmp, f"ok{rank}"), "w").write("1" if ok else "0")
    dist.destroy_process_group()


def test_sharded_step_matches_full_batch_gloo_world2(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, 31, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"ok{r}").read() == "1"
