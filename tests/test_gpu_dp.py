"""-m gpu, needs >= 2 GPUs (skipped otherwise; run with `gpurun --gpus 2`): data-parallel equivalence on hardware.
One optimisation step of a world-2 NCCL job on the two halves of a batch equals the world-1 step on the whole batch -
loss, flat gradient (after the bucketed, overlapped all-reduce), embedding gradient and the parameters after Adam - with the
labels skewed so that the per-rank counts of car / ignored / masked rays differ (the loss means must use GLOBAL counts:
SURVEY 8e, semantic/components/loss.py:35-65,117-157)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

N, ND, C, S = 2048, 512, 6, 64


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _global_batch():
    import numpy as np
    from oracle import render_oracle as O
    rng = np.random.Generator(np.random.PCG64(5))
    rays, extras = O.synthetic_rays(N, seed=40)
    lab = torch.from_numpy(rng.integers(0, C, (N, 1))).to(torch.uint8)
    lab[:300] = 4                     # every car-labelled ray of the first 300 sits in rank 0's shard ...
    lab[N // 2:][lab[N // 2:] == 4] = 1   # ... and rank 1 has none at all
    mask = torch.from_numpy(rng.uniform(0, 1, N) < 0.8)
    mask[N // 2: N // 2 + 600] = False    # rank 1 has far fewer rays in the cross-entropy mean
    batch = {"rays": rays, "extras": extras, "rgbs": torch.from_numpy(rng.uniform(0, 1, (N, 3))).float(), "semantic": lab,
             "semantic_sparsity_mask": mask}
    dr, de = O.synthetic_rays(ND, seed=41)
    depth = {"rays": dr, "extras": de, "depths": torch.from_numpy(rng.uniform(0.1, 0.5, (ND, 1))).float(),
             "weights": torch.from_numpy(rng.uniform(0, 1, (ND, 1))).float()}
    return batch, depth


def _make_trainer(dev, world, rank):
    from oracle import render_oracle as O
    from semnerf_b200.trainer import Trainer, default_cfgs
    cfgs = default_cfgs("semantic", n_samples=S, sc_lambda=0.05, use_car_reg_loss=True, car_reg_loss_start=0)
    tr = Trainer(cfgs, "semantic", C, device=dev, car_index=4, world=world, rank=rank, seed=0)
    params, emb = O.make_params(O.ModelSpec(kind="semantic", n_classes=C), seed=8)
    tr.models["coarse"].load_state_dict(params)
    tr.models["t"].weight.data.copy_(emb)
    return tr


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from semnerf_b200 import dist as snb_dist
    snb_dist.init_from_env(backend="nccl")
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    batch, depth = _global_batch()
    lo, hi = snb_dist.shard_range(N, rank, world)
    dlo, dhi = snb_dist.shard_range(ND, rank, world)
    shard = {k: v[lo:hi].to(dev) for k, v in batch.items()}
    dshard = {k: v[dlo:dhi].to(dev) for k, v in depth.items()}
    tr = _make_trainer(dev, world, rank)
    losses = []
    for step in range(2):
        loss = tr.training_step(shard, epoch=3, depth_batch=dshard, ray_offset=lo, global_rays=N, global_depth_rays=ND,
                                depth_ray_offset=dlo)
        torch.distributed.all_reduce(loss)          # the per-rank losses are shares of the global loss
        losses.append(loss.item())
        if step == 0:
            g0 = tr.gbuf.clone()
    ok = True
    msg = ""
    if rank == 0:
        ref = _make_trainer(dev, 1, 0)
        full = {k: v.to(dev) for k, v in batch.items()}
        dfull = {k: v.to(dev) for k, v in depth.items()}
        rl = []
        for step in range(2):
            rl.append(ref.training_step(full, epoch=3, depth_batch=dfull).item())
            if step == 0:
                gr = ref.gbuf.clone()
        a, b = g0.double(), gr.double()
        cos = float(a @ b) / float(a.norm() * b.norm())
        rel = float((a - b).norm() / b.norm())
        e_cos = float(a[:200] @ b[:200]) / float(a[:200].norm() * b[:200].norm())
        dp = float((tr.pbuf - ref.pbuf).abs().float().quantile(0.999))
        msg = (f"loss dp {losses} vs single {rl}; grad cosine {cos:.9f}, relative difference {rel:.3e}, embedding cosine "
               f"{e_cos:.7f}, parameters after 2 steps differ by {dp:.2e} (99.9th pct)")
        ok = (abs(losses[0] - rl[0]) <= 2e-5 * abs(rl[0]) and abs(losses[1] - rl[1]) <= 1e-3 * abs(rl[1])
              and cos >= 0.99999 and rel <= 5e-3 and e_cos >= 0.9999 and dp <= 1e-4)
    torch.distributed.barrier()
    open(os.path.join(tmp, f"r{rank}"), "w").write(("1 " if ok else "0 ") + msg)
    torch.distributed.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_world2_nccl_step_equals_world1_on_the_concatenated_batch(tmp_path):
    from semnerf_b200 import build
    build.build()
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = open(tmp_path / f"r{r}").read()
        print(res)
        assert res.startswith("1"), res
