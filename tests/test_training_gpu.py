"""The native training step (humanrf_b200.training.FusedTrainer): the loss goes down on a fixed batch, and its
flat gradient bucket equals what the autograd-compatible route (render + loss.backward()) puts in p.grad."""
import numpy as np
import pytest
import torch

from humanrf_b200.synthetic import make_model, synthetic_rays

pytestmark = pytest.mark.gpu


def test_loss_decreases_and_grads_match_autograd_route(cuda):
    from humanrf_b200.synthetic import input_batch_of
    from humanrf_b200.training import FusedTrainer
    from humanrf_b200.volume_rendering import render
    from oracle import rendering as R

    model, frames = make_model((6,), table_std=0.5, device=cuda)
    b = synthetic_rays(512, 64, frames, seed=4)
    g = {k: v.to(cuda).contiguous() for k, v in b.items() if k in ("o", "d", "frames", "t", "ri", "rgba")}
    tr = FusedTrainer(model, lr=1e-2, prune=False, seed=7)
    tr.keep_grad = True           # (the optimiser kernel normally leaves the bucket zeroed for the next step)
    # --- gradient equivalence on the first step (same background: replay the trainer's generator)
    gen = torch.Generator(device=cuda).manual_seed(7)
    bg = torch.rand((512, 3), device=cuda, generator=gen)
    ib = input_batch_of(b, cuda)
    out = render(ib, model, bg, is_training=True)
    loss, _ = R.training_loss(out.color, out.weights_sum, ib.rgba, bg)
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.hot_parameters()]).clone()
    before = [p.detach().clone() for p in model.hot_parameters()]
    l0 = tr.step(g["o"], g["d"], g["frames"], g["t"], g["ri"], g["rgba"], 512, return_loss=True)
    assert abs(l0 - loss.item()) < 1e-5 * max(1, abs(l0))
    rel = (tr.grad - ref).norm() / ref.norm()
    print("flat bucket vs autograd route relnorm", rel.item())
    assert rel < 1e-3          # identical kernels; only atomic ordering differs
    moved = sum(float((p.detach() - q).abs().max()) for p, q in zip(model.hot_parameters(), before))
    assert moved > 0
    losses = [l0] + [tr.step(g["o"], g["d"], g["frames"], g["t"], g["ri"], g["rgba"], 512, return_loss=True) for _ in range(40)]
    print("loss", losses[0], "->", losses[-1])
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < 0.7 * np.mean(losses[:5])


def test_prune_path_runs(cuda):
    from humanrf_b200.training import FusedTrainer

    model, frames = make_model((6, 6), table_std=0.5, device=cuda)
    b = synthetic_rays(256, 128, frames, seed=5, ragged=True)
    g = {k: v.to(cuda).contiguous() for k, v in b.items() if k in ("o", "d", "frames", "t", "ri", "rgba")}
    tr = FusedTrainer(model, prune=True)
    for _ in range(3):
        loss = tr.step(g["o"], g["d"], g["frames"], g["t"], g["ri"], g["rgba"], 256, return_loss=True)
        assert np.isfinite(loss)
    assert 0 < tr.last["samples"] <= g["t"].shape[0]


def test_int32_ray_indices_train_like_int64(cuda):
    """FusedTrainer.step accepts the ray index of a sample as int32 (a third of the upload of a host batch) and widens it on
    the device: the same seeds give the same losses as with the reference's int64 indices."""
    from humanrf_b200.training import FusedTrainer

    losses = []
    for dt in (torch.int64, torch.int32):
        model, frames = make_model((6,), table_std=0.5, device=cuda)
        b = synthetic_rays(256, 128, frames, seed=5, ragged=True)
        g = {k: v.to(cuda).contiguous() for k, v in b.items() if k in ("o", "d", "frames", "t", "ri", "rgba")}
        tr = FusedTrainer(model, prune=True, seed=3)
        losses.append([tr.step(g["o"], g["d"], g["frames"], g["t"], g["ri"].to(dt), g["rgba"], 256, return_loss=True) for _ in range(3)])
    assert losses[0][0] == losses[1][0], losses                      # first step: forward only, deterministic
    assert np.allclose(losses[0], losses[1], rtol=1e-4), losses      # later steps: fp32 atomics sum in any order


def test_multi_segment_steps_follow_the_reference_optimizer(cuda):
    """Three temporal segments, batches that touch only some of them.  The reference runs only the touched segments
    (humanrf.py:162-179), zero_grad(set_to_none=True) leaves the others' .grad at None (trainer.py:174) and
    torch.optim.Adam then skips them, per-parameter step counters included.  Checked: the autograd route hands None to
    the untouched segments; FusedTrainer leaves them bit-identical and counts their steps separately; after a few
    mixed steps both routes agree with each other."""
    from humanrf_b200.synthetic import input_batch_of
    from humanrf_b200.training import FusedTrainer
    from humanrf_b200.volume_rendering import render
    from oracle import rendering as R

    sizes = (6, 6, 6)
    ma, frames = make_model(sizes, table_std=0.5, device=cuda)
    mb, _ = make_model(sizes, table_std=0.5, device=cuda)
    for p, q in zip(ma.hot_parameters(), mb.hot_parameters()):
        assert torch.equal(p, q)
    init = [p.detach().clone() for p in ma.hot_parameters()]
    seg_frames = [frames[0:6], frames[6:12], frames[12:18]]
    plans = [(0, 2), (1,), (0, 1, 2), (2,)]                       # segments each step's batch draws its frames from
    opt = torch.optim.Adam(ma.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15)       # run.py:101
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda k: 0.5 ** min(k / 50001, 1))
    tr = FusedTrainer(mb, lr=1e-2, prune=False)
    counts = [0, 0, 0]
    for step, segs in enumerate(plans):
        fr = tuple(f for s_ in segs for f in seg_frames[s_])
        b = synthetic_rays(256, 48, fr, seed=30 + step, n_distinct_frames=len(fr))
        touched = sorted({int(ma.frame_numbers_to_segment_numbers[f]) for f in b["frames"].tolist()})
        assert set(touched) <= set(segs) and touched, touched
        for s_ in touched:
            counts[s_] += 1
        bg = torch.rand(256, 3, generator=torch.Generator().manual_seed(step)).to(cuda)
        # --- the reference's loop on the autograd route
        opt.zero_grad(set_to_none=True)
        ib = input_batch_of(b, cuda)
        out = render(ib, ma, bg, is_training=True)
        loss, _ = R.training_loss(out.color, out.weights_sum, ib.rgba, bg)
        loss.backward()
        for s_ in range(3):
            got_none = [p.grad is None for p in ma.hot_parameters()[5 * s_:5 * s_ + 5]]
            assert got_none == [s_ not in touched] * 5, (step, s_, got_none)
        assert all(p.grad is not None for p in ma.hot_parameters()[15:])
        opt.step()
        sched.step()
        # --- the fused trainer
        g = {k: v.to(cuda).contiguous() for k, v in b.items() if k in ("o", "d", "frames", "t", "ri", "rgba")}
        before = [p.detach().clone() for p in mb.hot_parameters()]
        lf = tr.step(g["o"], g["d"], g["frames"], g["t"], g["ri"], g["rgba"], 256, return_loss=True, background=bg)
        assert abs(lf - loss.item()) < 1e-3 * max(1.0, abs(lf)), (step, lf, loss.item())
        for s_ in range(3):
            same = [torch.equal(p.detach(), q) for p, q in zip(mb.hot_parameters()[5 * s_:5 * s_ + 5], before[5 * s_:5 * s_ + 5])]
            assert same == [s_ not in touched] * 5, (step, s_, same)
    assert tr.steps == [c for c in counts for _ in range(5)] + [len(plans)] * 2 and tr.t == len(plans)
    for i, (p, q, p0) in enumerate(zip(ma.hot_parameters(), mb.hot_parameters(), init)):
        moved = (p.detach() - p0).norm().item()
        rel = (p.detach() - q.detach()).norm().item() / max(moved, 1e-12)
        outliers = ((p.detach() - q.detach()).abs() > 2e-3).float().mean().item()
        print(f"param {i}: moved {moved:.3e} rel diff {rel:.3e} outliers {outliers:.2e}")
        assert rel < 2e-2 and outliers < 2e-3, (i, rel, outliers)


def test_train_loss_kernel_matches_the_reference_loss(cuda):
    """hrf_train_loss (forward + backward in one launch) vs trainer.py:205-215,229-238 through torch autograd, including
    rays whose weights_sum sits outside [0,1] (clamp) and residuals on both sides of the Huber delta."""
    from humanrf_b200 import _lib as L
    from oracle import rendering as R

    g = torch.Generator().manual_seed(3)
    n = 1000
    color = (torch.rand(n, 3, generator=g) * 1.2 - 0.1).to(cuda).requires_grad_(True)
    wsum = (torch.rand(n, 1, generator=g) * 1.4 - 0.2).to(cuda).requires_grad_(True)
    rgba = torch.rand(n, 4, generator=g)
    rgba[:, 3] = (rgba[:, 3] > 0.5).float()
    rgba[::5, :3] = (color.detach().cpu()[::5] + 0.004 * torch.randn(n, 3, generator=g)[::5]).clamp(0, 1)   # |x| < delta cases
    rgba, bg = rgba.to(cuda), torch.rand(n, 3, generator=g).to(cuda)
    loss, _ = R.training_loss(color, wsum, rgba, bg)
    (loss * 0.75).backward()
    scale = torch.tensor([0.75], device=cuda)
    dc, dw, out = torch.empty(n, 3, device=cuda), torch.empty(n, device=cuda), torch.zeros(1, device=cuda)
    L.check(L.lib().hrf_train_loss(color.data_ptr(), wsum.data_ptr(), rgba.data_ptr(), bg.data_ptr(), n, 0.01, 1e-3,
                                   scale.data_ptr(), dc.data_ptr(), dw.data_ptr(), out.data_ptr(), L.stream()))
    assert abs(out.item() - loss.item()) < 1e-6 * max(1.0, abs(loss.item()))
    torch.testing.assert_close(dc, color.grad, rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(dw, wsum.grad.view(-1), rtol=1e-5, atol=1e-9)
    assert (dw == 0).any() and (dw != 0).any()


def test_multi_tensor_adam_matches_torch_and_skips_inactive(cuda):
    """hrf_adam_multi: one launch over tensors of odd sizes / alignments == torch.optim.Adam per tensor; a tensor whose
    active flag is 0 keeps parameters, moments and its step counter; gradients are zeroed on request; the MLP-style
    permutation writes the bf16 copy through blob_perm."""
    import ctypes as C

    from humanrf_b200 import _lib as L

    g = torch.Generator().manual_seed(1)
    sizes = [8192 + 16, 4099, 24, 4096 * 3]
    ps = [torch.randn(n, generator=g).to(cuda) for n in sizes]
    refs = [p.clone().requires_grad_(True) for p in ps]
    opts = [torch.optim.Adam([r], lr=1e-2, betas=(0.9, 0.99), eps=1e-15) for r in refs]
    total = sum(sizes)
    flat_g, flat_m, flat_v = (torch.zeros(total, device=cuda) for _ in range(3))
    shadows = [torch.zeros(n, dtype=torch.bfloat16, device=cuda) for n in sizes]
    perm = torch.randperm(sizes[2], generator=g).to(torch.int32).to(cuda)
    active = torch.ones(len(sizes), dtype=torch.int32, device=cuda)
    steps = torch.zeros(len(sizes), dtype=torch.int32, device=cuda)
    items, first, off = [], 0, 0
    for i, n in enumerate(sizes):
        t = L.AdamTensor()
        t.param, t.exp_avg, t.exp_avg_sq, t.grad = ps[i].data_ptr(), flat_m[off:].data_ptr(), flat_v[off:].data_ptr(), flat_g[off:].data_ptr()
        t.shadow_bf16, t.blob_perm = shadows[i].data_ptr(), (perm.data_ptr() if i == 2 else None)
        t.active, t.step, t.n, t.first_block = active[i:].data_ptr(), steps[i:].data_ptr(), n, first
        first += (n + L.ADAM_BLOCK_ELEMS - 1) // L.ADAM_BLOCK_ELEMS
        off += n
        items.append(t)
    desc = torch.from_numpy(np.frombuffer(b"".join(bytes(x) for x in items), dtype=np.uint8).copy()).to(cuda)
    plan = [(1, 1, 1, 1), (1, 0, 1, 1), (1, 1, 1, 0), (1, 1, 1, 1)]
    for it, act in enumerate(plan):
        active.copy_(torch.tensor(act, dtype=torch.int32))
        off = 0
        for i, n in enumerate(sizes):
            gr = (torch.randn(n, generator=g) * 10 ** float(torch.randint(-5, 1, (1,), generator=g))).to(cuda)
            flat_g[off:off + n] = gr
            if act[i]:
                refs[i].grad = gr.clone()
                opts[i].step()
            off += n
        L.check(L.lib().hrf_adam_multi(desc.data_ptr(), len(sizes), first, 1e-2, 0.9, 0.99, 1e-15, 1.0, 1, L.stream()))
        for i in range(len(sizes)):
            np.testing.assert_allclose(ps[i].cpu().numpy(), refs[i].detach().cpu().numpy(), rtol=3e-6, atol=3e-7)
        off = 0
        for i, n in enumerate(sizes):               # active tensors' gradients are cleared, inactive ones are not touched
            assert bool((flat_g[off:off + n] == 0).all()) == bool(act[i])
            off += n
    assert steps.cpu().tolist() == [sum(a[i] for a in plan) for i in range(len(sizes))]
    for i, n in enumerate(sizes):
        want = ps[i].to(torch.bfloat16)
        if i == 2:
            got = shadows[i][perm.long()]
        else:
            got = shadows[i]
        torch.testing.assert_close(got.float(), want.float(), rtol=0, atol=0)


@pytest.mark.parametrize("segs", [(6,), (6, 6)], ids=["1seg", "2seg"])
def test_feature_reuse_modes_train_alike(cuda, segs):
    """reuse="feat" (the survivors' forward runs the MLPs on the prune pass's composed features) and "feat+grid" (the
    scatter also takes the prune pass's per-grid features) against "none" (prune pass, then a full forward of the
    survivors): same jitter, same survivors, same loss to float rounding, parameters that move the same way."""
    from humanrf_b200.training import FusedTrainer

    b = None
    results = {}
    for reuse in ("none", "feat", "feat+grid"):
        model, frames = make_model(segs, table_std=3.0, device=cuda)      # dense enough for rays to saturate: pruning bites
        if b is None:
            b = synthetic_rays(300, 400, frames, seed=6, ragged=True)        # long rays: they saturate, pruning removes samples
        g = {k: v.to(cuda).contiguous() for k, v in b.items() if k in ("o", "d", "frames", "t", "ri", "rgba")}
        init = [p.detach().clone() for p in model.hot_parameters()]
        tr = FusedTrainer(model, lr=1e-2, prune=True, seed=11, reuse=reuse)
        losses, kept = [], []
        for _ in range(3):
            losses.append(tr.step(g["o"], g["d"], g["frames"], g["t"], g["ri"], g["rgba"], 300, return_loss=True))
            kept.append(int(tr.last["samples"]))
        results[reuse] = (losses, kept, [p.detach().clone() for p in model.hot_parameters()], init)
    l0, k0, p0, init = results["none"]
    assert 0 < k0[0] < g["t"].shape[0]
    for reuse in ("feat", "feat+grid"):
        l1, k1, p1, _ = results[reuse]
        assert k1[0] == k0[0], (reuse, k0, k1)                   # first step: identical parameters -> identical pruning
        assert abs(l1[0] - l0[0]) < 1e-6 * max(1.0, abs(l0[0])), (reuse, l0, l1)
        np.testing.assert_allclose(l1, l0, rtol=2e-2)
        for i, (a, c, s0) in enumerate(zip(p1, p0, init)):
            moved = (c - s0).norm().item()
            rel = (a - c).norm().item() / max(moved, 1e-12)
            assert rel < 5e-2, (reuse, i, rel)
