"""Multi-GPU paths: (1) 2-rank NCCL data-parallel training reproduces single-process training on the union batch;
(2) tile-sharded inference assembles the same image as one monolithic pass (no collective)."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from humanrf_b200.synthetic import make_model, synthetic_rays
from scene import make_scene

pytestmark = pytest.mark.gpu
STEPS = 3


def _subset(b, lo, hi):
    sel = (b["ri"] >= lo) & (b["ri"] < hi)
    return dict(o=b["o"][lo:hi], d=b["d"][lo:hi], frames=b["frames"][lo:hi], rgba=b["rgba"][lo:hi], t=b["t"][sel],
                ri=b["ri"][sel] - lo)


def _train(rank, world, port, spans, q):
    import torch.distributed as dist

    from humanrf_b200.training import FusedTrainer

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    pg = None
    if world > 1:
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    model, frames = make_model((6,), table_std=0.5, device=dev)
    b = synthetic_rays(512, 48, frames, seed=4, ragged=True)
    bg_all = torch.rand(512, 3, generator=torch.Generator().manual_seed(9))
    lo, hi = spans[rank]
    sb = {k: v.to(dev).contiguous() for k, v in _subset(b, lo, hi).items()}
    tr = FusedTrainer(model, lr=1e-2, prune=False, world_size=world)
    for _ in range(STEPS):
        tr.step(sb["o"], sb["d"], sb["frames"], sb["t"], sb["ri"], sb["rgba"], hi - lo, background=bg_all[lo:hi].to(dev))
    torch.cuda.synchronize()
    if rank == 0:
        q.put([p.detach().cpu().numpy() for p in model.hot_parameters()])
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_two_rank_dp_equals_single_process(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_train, args=(0, 1, 0, [(0, 512)], q))
    p.start(); ref = q.get(timeout=300); p.join()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_train, args=(r, 2, port, [(0, 200), (200, 512)], q)) for r in range(2)]   # unequal shares
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    for a, b in zip(got, ref):
        moved = np.abs(b).max()
        assert np.abs(a - b).max() <= 2e-3 * moved + 1e-6, (np.abs(a - b).max(), moved)


def test_tile_sharded_render_equals_monolithic(cuda):
    from humanrf_b200.dataset.occupancy_grid_native import OccupanyGrid
    from humanrf_b200.parallel import TileShardedRenderer, shard_range

    model, frames = make_model((6,), table_std=0.5, device=cuda)
    sc = make_scene(num_images=1, width=160, height=120, G=64, portrait_every=0)
    og = OccupanyGrid(sc["G"], 1)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    cam = dict(frame_numbers=t(sc["frame_numbers"]), camera_numbers=t(sc["camera_numbers"]),
               grid_handles=torch.tensor([og.add_grid(t(sc["grids"][0]))], dtype=torch.int64, device=cuda),
               landscape=t(sc["landscape"]), inverse_krs=t(sc["inverse_krs"]), camera_origins=t(sc["camera_origins"]),
               aabb=t(sc["aabb"]), G=sc["G"], width=160, height=120)
    r = TileShardedRenderer(model, og, rays_per_batch=4096)
    full = r.render_range(cam, 0, 160 * 120)
    parts = []
    for rank in range(3):
        s, e, c = r.render_image_sharded(cam, rank, 3)
        assert (s, e) == shard_range(160 * 120, rank, 3)
        parts.append(c)
    torch.testing.assert_close(torch.cat(parts), full, rtol=0, atol=0)
    assert full.abs().sum() > 0 and (full == 0).any()          # object pixels rendered, background left at 0
