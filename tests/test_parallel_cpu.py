"""World-size-2 gloo tests (CPU) of the data-parallel host logic: union-batch loss weighting + bucket all-reduce
reproduce the single-process gradient; tile / sequence sharding covers every item exactly once."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from humanrf_b200.parallel import (active_segments, allreduce_bucket_, allreduce_spans, broadcast_parameters_, deal_round_robin,
                                   grid_major_bucket_layout, mask_inactive_segment_grads, shard_range, union_batch_loss_scale)


def test_shard_range_and_round_robin_cover_everything():
    for n in (0, 1, 7, 773056):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1
    seq = [(c, f) for c in range(5) for f in range(7)]
    dealt = [deal_round_robin(seq, r, 4) for r in range(4)]
    assert sorted(sum(dealt, [])) == sorted(seq)


def test_grid_major_bucket_layout():
    for S, extra in ((1, 2), (2, 2), (3, 3), (5, 2)):
        sizes = []
        for s in range(S):
            sizes += [1000 + 10 * s + k for k in range(4)] + [77 + s]           # 4 grids + vectors per segment
        sizes += [3072, 7168, 320][:extra]                                      # sigma net, colour net, [embeddings]
        slices, regions, order = grid_major_bucket_layout(sizes, S)
        assert sorted(order) == list(range(len(sizes))) and len(regions) == 5
        # the slices tile the bucket exactly, in `order`
        pos = 0
        for i in order:
            assert slices[i] == (pos, pos + sizes[i])
            pos = slices[i][1]
        assert regions[0][0] == 0 and regions[-1][1] == pos and all(a[1] == b[0] for a, b in zip(regions, regions[1:]))
        for k in range(4):                                                      # region k = grid k of every segment, nothing else
            inside = [i for i in range(len(sizes)) if regions[k][0] <= slices[i][0] and slices[i][1] <= regions[k][1]]
            assert inside == [5 * s + k for s in range(S)]
        # sigma net and colour net stay adjacent (the fused backward writes them as one block)
        assert slices[5 * S][1] == slices[5 * S + 1][0]


def test_allreduce_spans_cover_exactly_the_active_gradients():
    for S in (1, 2, 4):
        sizes = []
        for s in range(S):
            sizes += [100 + s, 200 + s, 300 + s, 400 + s, 50 + s]
        sizes += [3072, 7168, 320]
        slices, regions, _ = grid_major_bucket_layout(sizes, S)
        # nothing to decide: one span per region, together the whole bucket
        assert allreduce_spans(slices, regions, S, None) == [[r] for r in regions]
        assert allreduce_spans(slices, regions, S, [True] * S) == [[r] for r in regions]
        for active in ([False] * S, [s % 2 == 0 for s in range(S)], [s == S - 1 for s in range(S)]):
            spans = allreduce_spans(slices, regions, S, active)
            covered = torch.zeros(regions[-1][1], dtype=torch.int32)
            for k, region in enumerate(spans):
                for a, b in region:
                    assert regions[k][0] <= a < b <= regions[k][1]        # spans stay inside their region
                    covered[a:b] += 1
            want = torch.zeros_like(covered)
            for i, (a, b) in enumerate(slices):
                if i >= 5 * S or active[i // 5]:
                    want[a:b] = 1
            assert torch.equal(covered, want)                              # every active gradient once, nothing else
            assert all(x[1] < y[0] for region in spans for x, y in zip(region, region[1:]))   # merged: no touching spans


def test_active_segments_and_gradient_masking():
    """humanrf.py:162-179 + trainer.py:174: segments the batch's frames do not fall into take no part in the step."""
    from oracle.field import frame_luts

    frames = tuple(range(15, 15 + 6 + 12 + 6))
    f2s, _ = frame_luts(frames, (6, 12, 6))
    lut = torch.from_numpy(f2s)
    assert active_segments(lut, torch.tensor([15, 16, 20]), 3).tolist() == [True, False, False]
    assert active_segments(lut, torch.tensor([[33], [21]], dtype=torch.int32), 3).tolist() == [False, True, True]
    assert active_segments(lut, torch.tensor([3, 14, 999, -1]), 3).tolist() == [False, False, False]   # no segment / out of range
    assert active_segments(lut, torch.zeros(0, dtype=torch.int64), 3).tolist() == [False, False, False]
    grads = list(range(5 * 3 + 2))
    masked = mask_inactive_segment_grads(grads, [True, False, True])
    assert masked[:5] == grads[:5] and masked[5:10] == [None] * 5 and masked[10:] == grads[10:]


def test_optimiser_shards_partition_every_tensor_exactly_once():
    """FusedTrainer exchange="p2p": rank r owns shard_bounds(n, r, world) of every hash-table tensor (reduce-scatter +
    sharded Adam + all-gather of the bf16 shadows in hrf_dp_reduce_adam).  The slices must tile [0, n) without overlap,
    start on the Adam kernel's block boundary, and degrade to empty slices for tensors smaller than the world."""
    from humanrf_b200.parallel import shard_bounds

    for n in (0, 1, 4095, 4096, 4097, 7_391_536, 13_969_152, 3 * 4096 * 8, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            cover = 0
            prev_end = 0
            for r in range(world):
                a, b = shard_bounds(n, r, world, 4096)
                assert 0 <= a <= b <= n and a == prev_end if b > a else a <= n
                if b > a:
                    assert a % 4096 == 0
                    cover += b - a
                    prev_end = b
            assert cover == n
            sizes = [shard_bounds(n, r, world, 4096) for r in range(world)]
            lens = [b - a for a, b in sizes]
            assert max(lens) - min(l for l in lens if l or True) <= max(lens)      # (monotone non-increasing)
            assert lens == sorted(lens, reverse=True)


def test_allreduce_helpers_are_no_ops_without_a_process_group():
    t = torch.arange(4.0)
    assert allreduce_bucket_(t) is t and allreduce_bucket_(t, async_op=True) is None
    assert union_batch_loss_scale(7, "cpu") == 1.0


def _worker(rank, world, port, ray_counts, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    w = torch.randn(16, requires_grad=True)                     # replicated "parameters"
    g = torch.Generator().manual_seed(1)
    x = torch.randn(sum(ray_counts), 16, generator=g)            # the union batch, identical on every rank
    y = torch.randn(sum(ray_counts), generator=g)
    lo = sum(ray_counts[:rank])
    xs, ys = x[lo:lo + ray_counts[rank]], y[lo:lo + ray_counts[rank]]
    loss = torch.nn.functional.huber_loss(xs @ w, ys, delta=0.01, reduction="mean")   # per-rank mean, as FusedTrainer
    (loss * union_batch_loss_scale(ray_counts[rank], "cpu")).backward()
    bucket = w.grad.clone()
    regions = w.grad.clone()                                     # the same bucket reduced region by region, asynchronously
    allreduce_bucket_(bucket)
    works = [allreduce_bucket_(regions[a:b], async_op=True) for a, b in ((0, 5), (5, 6), (6, 16))]
    for wk in works:
        wk.wait()
    assert torch.equal(regions, bucket)                          # FusedTrainer's overlapped schedule == one message
    bucket /= world                                              # Adam's grad_scale = 1/world
    p = torch.full((4,), float(rank))
    broadcast_parameters_([p], src=0)
    if rank == 0:
        w2 = w.detach().clone().requires_grad_(True)
        torch.nn.functional.huber_loss(x @ w2, y, delta=0.01, reduction="mean").backward()
        out.put((bucket.numpy(), w2.grad.numpy()))
    assert float(p.sum()) == 0.0
    dist.barrier()
    dist.destroy_process_group()


def test_union_batch_weighting_matches_single_process_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, (37, 91), out)) for r in range(2)]
    for p in procs:
        p.start()
    got, ref = out.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-7)
