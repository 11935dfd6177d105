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


def _collect(q, procs, count, timeout=240):
    """`count` results from the queue, failing at once if a child died instead of waiting out the timeout."""
    import queue
    import time

    out, t0 = [], time.time()
    while len(out) < count:
        try:
            out.append(q.get(timeout=2))
        except queue.Empty:
            dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
            assert not dead, f"a rank exited with {dead}"
            assert time.time() - t0 < timeout, "timed out waiting for the ranks"
    return out


def _train(rank, world, port, spans, q, exchange="p2p", segs=(6,), frame_plan=None):
    """frame_plan: per step, the frames the batch's rays are re-drawn from (multi-segment case: steps that touch only
    some segments, and different segments on different ranks)."""
    import torch.distributed as dist

    from humanrf_b200.training import FusedTrainer

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    if world > 1:
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    model, frames = make_model(segs, table_std=0.5, device=dev)
    b = synthetic_rays(512, 48, frames, seed=4, ragged=True)
    bg_all = torch.rand(512, 3, generator=torch.Generator().manual_seed(9))
    lo, hi = spans[rank]
    tr = FusedTrainer(model, lr=1e-2, prune=False, world_size=world, exchange=exchange.split("-")[0],
                      overlap_exchange=exchange.endswith("-overlap"))
    for step in range(STEPS):
        if frame_plan is not None:
            pool = torch.tensor(frame_plan[step], dtype=torch.int32)
            h = pool.numel() // 2                                        # rays < 200 (rank 0's share) draw from the first half of
            r_ = torch.arange(512)                                       # the pool, the others from the second: the two ranks
            b["frames"] = torch.where(r_ < 200, pool[r_ % h], pool[h + r_ % (pool.numel() - h)])   # may touch DIFFERENT segments
        sb = {k: v.to(dev).contiguous() for k, v in _subset(b, lo, hi).items()}
        tr.step(sb["o"], sb["d"], sb["frames"], sb["t"], sb["ri"], sb["rgba"], hi - lo, background=bg_all[lo:hi].to(dev))
    tr.gather_master_parameters()      # exchange="p2p" shards the fp32 masters of the tables over the ranks
    torch.cuda.synchronize()
    shadows = [s.float().cpu().numpy() for row in model.native().shadows for s in row]
    q.put((rank, [p.detach().cpu().numpy() for p in model.hot_parameters()], shadows, tr.steps))
    if world > 1:
        dist.barrier()
        tr.close()
        dist.destroy_process_group()


def _compare_with_single_process(got, ref, init):
    # the replicas apply the same reduced gradient: they must stay bit-identical (masters after the gather, and the
    # bf16 shadow tables every rank's forward actually reads)
    diverged = [(i, float(np.abs(a - b).max())) for i, (a, b) in enumerate(zip(got[0][0], got[1][0])) if not np.array_equal(a, b)]
    diverged += [("shadow", i) for i, (a, b) in enumerate(zip(got[0][1], got[1][1])) if not np.array_equal(a, b)]
    print("replica divergence:", diverged)
    assert got[0][2] == got[1][2] == ref[2], (got[0][2], ref[2])          # per-parameter Adam step counters
    # Against single-process training on the union batch.  Adam (eps = 1e-15) moves an entry by ~lr whatever the size of
    # its gradient, so an entry whose contributions cancel to rounding noise may step the other way: compare the bulk
    # (relative L2 of the difference against the distance trained) and bound the share of such outliers.
    bad = []
    for i, (a, b, p0) in enumerate(zip(got[0][0], ref[0], init)):
        diff, moved = np.abs(a - b), np.linalg.norm(b - p0)
        rel = np.linalg.norm(a - b) / max(moved, 1e-12)
        outliers = float((diff > 2e-3).mean())
        print(f"param {i}: size {a.size} max diff {diff.max():.3e} rel L2 {rel:.3e} outliers {outliers:.2e}")
        if not (rel <= 1e-4 and outliers <= 1e-5) and moved > 0:      # measured: rel 2e-7 .. 2e-6, no outliers
            bad.append((i, rel, outliers, diff.max()))
    assert not diverged, diverged
    assert not bad, bad


def _run_dp_case(cuda, exchange, segs, frame_plan, port_offset):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    init = [p.detach().cpu().numpy() for p in make_model(segs, table_std=0.5, device=cuda)[0].hot_parameters()]
    p = ctx.Process(target=_train, args=(0, 1, 0, [(0, 512)], q, exchange, segs, frame_plan))
    p.start(); (_, *ref), = _collect(q, [p], 1); p.join(timeout=60)
    port = 29600 + os.getpid() % 1000 + port_offset
    procs = [ctx.Process(target=_train, args=(r, 2, port, [(0, 200), (200, 512)], q, exchange, segs, frame_plan)) for r in range(2)]   # unequal shares
    for p in procs:
        p.start()
    got = {r: rest for r, *rest in _collect(q, procs, 2)}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    _compare_with_single_process(got, ref, init)


EXCHANGES = ["nccl", "p2p", "p2p-overlap"]


@pytest.mark.parametrize("exchange", EXCHANGES)
def test_two_rank_dp_equals_single_process(cuda, exchange):
    """exchange="nccl": one all-reduce of the whole bucket, Adam on every rank; "p2p" (FusedTrainer's default): the
    reduce-scatter + rank-sharded Adam + all-gather of the bf16 shadows as ONE kernel over NVLink peer memory after the
    scatter; "p2p-overlap": the same kernel once per hash grid on a side stream while the next grid is scattered."""
    _run_dp_case(cuda, exchange, (6,), None, EXCHANGES.index(exchange))


@pytest.mark.parametrize("exchange", EXCHANGES)
def test_two_rank_dp_multi_segment(cuda, exchange):
    """Three temporal segments; steps whose batches touch only some of them: a segment takes part in the step (gradient
    exchange, Adam, step counter) iff ANY rank's batch touches it -- the union-batch semantics of the reference's
    optimiser (untouched segments keep .grad = None, trainer.py:174).  DESIGN.md 7.4 of round 1: never run on GPUs."""
    frames = list(range(15, 15 + 18))
    plan = [frames[0:3] + frames[12:15], frames[6:9], frames[0:18:3]]       # segments {0,2}, {1}, {0,1,2}
    assert STEPS == len(plan)
    _run_dp_case(cuda, exchange, (6, 6, 6), plan, 3 + EXCHANGES.index(exchange))


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
    # (the compositing runs per 128-sample tile inside the field kernel: another split of the pixel range shifts the tile
    #  borders, i.e. the summation order of a ray's weights -- float rounding, not bit identity)
    torch.testing.assert_close(torch.cat(parts), full, rtol=0, atol=2e-6)
    assert full.abs().sum() > 0 and (full == 0).any()          # object pixels rendered, background left at 0
