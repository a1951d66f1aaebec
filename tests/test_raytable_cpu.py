"""CPU tests of the device-resident ray table / batch sampler and the point-cloud helpers (host logic; the tables
work on any torch device)."""
import torch

from semnerf_b200.pointcloud import denormalize, xyz_from_depth
from semnerf_b200.raytable import DeviceRayTable, ZippedTables


def _table(n=1000):
    rays = torch.arange(n * 8, dtype=torch.float32).view(n, 8)
    return DeviceRayTable({"rays": rays, "ids": torch.arange(n)}, device="cpu")


def test_epoch_visits_every_ray_once_and_keeps_keys_aligned():
    t = _table(1000)
    seen = []
    for b in t.epoch(128, seed=3, epoch=0):
        assert torch.equal(b["rays"][:, 0], b["ids"].float() * 8)      # rows of different keys stay together
        seen.append(b["ids"])
    sizes = [len(s) for s in seen]
    assert sizes == [128] * 7 + [104]                                  # drop_last=False: short last batch
    allids = torch.cat(seen)
    assert torch.equal(torch.sort(allids).values, torch.arange(1000))
    assert not torch.equal(allids, torch.arange(1000))                 # shuffled
    again = torch.cat([b["ids"] for b in t.epoch(128, seed=3, epoch=0)])
    other = torch.cat([b["ids"] for b in t.epoch(128, seed=3, epoch=1)])
    assert torch.equal(again, allids) and not torch.equal(other, allids)


def test_rank_shards_partition_every_global_batch():
    t = _table(777)
    full = list(t.epoch(100, seed=5, epoch=2))
    for world in (2, 3, 8):
        shards = [list(t.epoch(100, seed=5, epoch=2, rank=r, world=world)) for r in range(world)]
        for step, b in enumerate(full):
            got = torch.cat([shards[r][step]["ids"] for r in range(world)])
            assert torch.equal(got, b["ids"])


def test_drop_last_and_no_shuffle():
    t = _table(250)
    bs = list(t.epoch(100, shuffle=False, drop_last=True))
    assert len(bs) == 2 and torch.equal(bs[0]["ids"], torch.arange(100)) and t.steps_per_epoch(100, True) == 2
    assert t.steps_per_epoch(100) == 3


def test_zipped_tables_cycle_the_shorter_one():
    z = ZippedTables({"color": _table(1000), "depth": _table(150)}, {"color": 100, "depth": 100}, seed=1)
    steps = list(z.epoch(0))
    assert len(steps) == 10 and all(set(s) == {"color", "depth"} for s in steps)
    assert [len(s["depth"]["ids"]) for s in steps[:4]] == [100, 50, 100, 50]


def test_pointcloud_helpers_match_reference_formulas():
    g = torch.Generator().manual_seed(0)
    rays = torch.rand(64, 8, generator=g)
    depth = torch.rand(64, generator=g)
    xyz = xyz_from_depth(rays, depth)
    assert xyz.dtype == torch.float64
    ref = rays[:, 0:3].double() + rays[:, 3:6].double() * depth.double().view(-1, 1)
    assert torch.equal(xyz, ref)
    out = denormalize(xyz, (10.0, 20.0, 30.0), 4.0)
    assert torch.allclose(out, ref * 4.0 + torch.tensor([10.0, 20.0, 30.0], dtype=torch.float64))
