"""Parity of the per-ray scan kernels (prune / weights / accumulate and their backward) with the oracle
restating nerfacc 0.3.1 as called from volume_rendering.py:75-84,123-145."""
import numpy as np
import pytest
import torch

from helpers import synthetic_rays
from humanrf_b200 import _lib as L
from humanrf_b200.volume_rendering import ray_offsets
from oracle import rendering as R

pytestmark = pytest.mark.gpu


def _batch(num_rays, spr, seed=1, ragged=True, sigma_scale=400.0):
    b = synthetic_rays(num_rays, spr, tuple(range(15, 21)), seed=seed, ragged=ragged)
    g = torch.Generator().manual_seed(seed)
    n = b["t"].shape[0]
    sigma = torch.rand(n, generator=g) ** 4 * sigma_scale
    sigma[torch.rand(n, generator=g) < 0.3] *= 1e-3      # plenty of alpha < 1e-4
    rgb = torch.rand(n, 3, generator=g)
    return b, sigma, rgb


def test_ray_offsets_and_prune(cuda):
    for seed, (nr, spr) in enumerate([(257, 70), (64, 513), (5, 3), (40, 0)]):
        b, sigma, _ = _batch(nr, max(spr, 1), seed, ragged=spr != 513)
        if spr == 0:
            b["ri"], b["t"], sigma = b["ri"][:0], b["t"][:0], sigma[:0]
        ri, t = b["ri"].to(cuda), b["t"].to(cuda)
        off = ray_offsets(ri, nr)
        exp_off = np.searchsorted(b["ri"].numpy(), np.arange(nr + 1))
        np.testing.assert_array_equal(off.cpu().numpy(), exp_off)
        n = t.shape[0]
        keep = torch.empty(n, dtype=torch.uint8, device=cuda)
        kept_off = torch.empty(nr + 1, dtype=torch.int32, device=cuda)
        out_t = torch.empty(n, device=cuda)
        out_ri = torch.empty(n, dtype=torch.int64, device=cuda)
        counter = torch.zeros(1, dtype=torch.int64, device=cuda)
        src = torch.empty(n, dtype=torch.int32, device=cuda)
        L.check(L.lib().hrf_prune(sigma.to(cuda).data_ptr(), t.data_ptr(), ri.data_ptr(), off.data_ptr(), nr, 4e-4, 1e-4,
                                  1e-4, keep.data_ptr(), kept_off.data_ptr(), out_t.data_ptr(), out_ri.data_ptr(),
                                  src.data_ptr(), counter.data_ptr(), L.stream()))
        exp = R.prune_mask(sigma, b["ri"]).numpy()
        got = keep.bool().cpu().numpy()
        # identical except where T or alpha sits within float rounding of the 1e-4 thresholds
        alphas = 1.0 - torch.exp(-sigma.double() * 4e-4)
        T = R._exclusive_by_ray((1.0 - alphas), b["ri"], "prod").numpy()
        near = (np.abs(T - 1e-4) < 1e-8) | (np.abs(alphas.numpy() - 1e-4) < 2.5e-7)   # alpha = 1 - exp(..) cancels to ~1e-7 abs
        assert (got != exp)[~near].sum() == 0, ((got != exp).sum(), near.sum())
        k = int(counter.item())
        assert k == got.sum()
        np.testing.assert_array_equal(out_t[:k].cpu().numpy(), b["t"].numpy()[got])
        np.testing.assert_array_equal(out_ri[:k].cpu().numpy(), b["ri"].numpy()[got])
        np.testing.assert_array_equal(src[:k].cpu().numpy(), np.nonzero(got)[0])          # where each survivor came from
        np.testing.assert_array_equal(kept_off.cpu().numpy(), np.concatenate(([0], np.cumsum(np.bincount(b["ri"].numpy()[got], minlength=nr)))))


@pytest.mark.parametrize("with_bg", [False, True])
def test_composite_forward_backward(cuda, with_bg):
    nr = 193
    b, sigma, rgb = _batch(nr, 90, seed=7)
    g = torch.Generator().manual_seed(3)
    bg = torch.rand(nr, 3, generator=g) if with_bg else None
    s64 = sigma.double().requires_grad_(True)
    c64 = rgb.double().requires_grad_(True)
    # oracle in float64 but with the float32 dt = (t+step)-t the reference feeds nerfacc
    t32 = b["t"]
    dt = ((t32 + 4e-4) - t32).double()
    sdt = s64 * dt
    T = torch.exp(-R._exclusive_by_ray(sdt, b["ri"], "sum"))
    w = T * (1 - torch.exp(-sdt))
    col = R.accumulate(w, b["ri"], c64, nr)
    ws = R.accumulate(w, b["ri"], None, nr)
    if with_bg:
        col = col + bg.double() * (1 - ws)
    dcol = torch.randn(nr, 3, generator=g).double()
    dws = torch.randn(nr, 1, generator=g).double()
    (col * dcol).sum().add((ws * dws).sum()).backward()

    off = ray_offsets(b["ri"].to(cuda), nr)
    sig_d, rgb_d, t_d = sigma.to(cuda), rgb.to(cuda).contiguous(), b["t"].to(cuda)
    color = torch.empty(nr, 3, device=cuda)
    wsum = torch.empty(nr, device=cuda)
    wts = torch.empty(t_d.shape[0], device=cuda)
    bg_d = bg.to(cuda).contiguous() if with_bg else None
    lib = L.lib()
    L.check(lib.hrf_composite_forward(sig_d.data_ptr(), rgb_d.data_ptr(), t_d.data_ptr(), off.data_ptr(), nr, 4e-4,
                                      L.ptr(bg_d), color.data_ptr(), wsum.data_ptr(), wts.data_ptr(), L.stream()))
    np.testing.assert_allclose(wts.cpu().numpy(), w.detach().numpy(), rtol=2e-4, atol=2e-7)
    np.testing.assert_allclose(color.cpu().numpy(), col.detach().numpy(), rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(wsum.cpu().numpy(), ws.detach().numpy()[:, 0], rtol=2e-4, atol=2e-6)
    # fp32 oracle path used by the rest of the suite agrees too
    c32, w32 = R.render(b["t"], sigma, rgb, b["ri"], nr, bg)
    np.testing.assert_allclose(color.cpu().numpy(), c32.numpy(), rtol=1e-3, atol=1e-5)

    d_sigma = torch.empty_like(sig_d)
    d_rgb = torch.empty_like(rgb_d)
    dc_d, dw_d = dcol.float().to(cuda).contiguous(), dws.float().reshape(-1).to(cuda).contiguous()
    L.check(lib.hrf_composite_backward(sig_d.data_ptr(), rgb_d.data_ptr(), t_d.data_ptr(), off.data_ptr(), nr, 4e-4,
                                       L.ptr(bg_d), dc_d.data_ptr(), dw_d.data_ptr(), d_sigma.data_ptr(),
                                       d_rgb.data_ptr(), L.stream()))
    gs, gc = s64.grad.numpy(), c64.grad.numpy()
    np.testing.assert_allclose(d_rgb.cpu().numpy(), gc, rtol=3e-4, atol=1e-6)
    np.testing.assert_allclose(d_sigma.cpu().numpy(), gs, rtol=2e-3, atol=2e-6 * np.abs(gs).max())


def test_early_stop_density_pass_gives_the_same_kept_set(cuda):
    """hrf_field_density_early_stop skips chunks behind an opaque prefix; hrf_prune on its sigma must keep exactly the
    samples it keeps on the fully evaluated sigma, and the evaluated densities must be bit-identical."""
    from helpers import make_pair

    om, m, frames = make_pair((6,), table_std=6.0)
    m.density_scale = 400.0                      # opaque quickly: T < 1e-4 after ~60 samples
    m._native = None
    nat = m.native()
    for ragged in (False, True):
        b = synthetic_rays(300, 700, frames, seed=5, ragged=ragged)
        g = {k: v.to(cuda).contiguous() for k, v in b.items() if k in ("o", "d", "frames", "t", "ri")}
        s = nat.samples_rays(g["o"], g["d"], g["frames"], g["t"], g["ri"])
        full, _, _, _ = nat.forward(s, 0, want_geo=False, want_feat=False)
        off = ray_offsets(g["ri"], 300)
        es = nat.density_early_stop(s, off, 300, 4e-4)
        torch.cuda.synchronize()
        skipped = (es == 0) & (full != 0)
        assert skipped.any(), "the schedule never skipped anything"
        torch.testing.assert_close(es[~skipped], full[~skipped], rtol=0, atol=0)
        masks = []
        for sig in (full, es):
            n = sig.shape[0]
            keep = torch.empty(n, dtype=torch.uint8, device=cuda)
            kept_off = torch.empty(301, dtype=torch.int32, device=cuda)
            ot, ori = torch.empty(n, device=cuda), torch.empty(n, dtype=torch.int64, device=cuda)
            cnt = torch.zeros(1, dtype=torch.int64, device=cuda)
            L.check(L.lib().hrf_prune(sig.data_ptr(), g["t"].data_ptr(), g["ri"].data_ptr(), off.data_ptr(), 300, 4e-4, 1e-4,
                                      1e-4, keep.data_ptr(), kept_off.data_ptr(), ot.data_ptr(), ori.data_ptr(),
                                      None, cnt.data_ptr(), L.stream()))
            masks.append(keep.clone())
        assert torch.equal(masks[0], masks[1])
        print(f"ragged={ragged}: {int(skipped.sum())} of {full.numel()} densities skipped, kept {int(masks[0].sum())}")


@pytest.mark.parametrize("with_bg", [False, True])
def test_fused_render_kernel_equals_forward_plus_composite(cuda, with_bg):
    """hrf_render_fused (compositing as the epilogue of the field kernel; rays cut by 128-sample tile borders chained by
    the fix-up kernel) vs hrf_field_forward + hrf_composite_forward on the same samples: rays shorter than a tile, rays
    spanning 2..5 tiles, empty rays, a tail tile; then with a live sample count below the capacity."""
    import ctypes as C

    from helpers import make_pair
    from humanrf_b200.volume_rendering import render_fused

    _, m, frames = make_pair((6,), table_std=2.0)
    nr = 150
    b = synthetic_rays(nr, 600, frames, seed=13, ragged=True)          # 0..600 samples per ray, every 7th ray empty
    g = {k: v.to(cuda).contiguous() for k, v in b.items() if k in ("o", "d", "frames", "t", "ri")}
    n = g["t"].shape[0]
    counts = torch.bincount(b["ri"], minlength=nr)
    assert (counts == 0).any() and (counts > 512).any() and (counts < 128).any()
    bg = torch.rand(nr, 3, generator=torch.Generator().manual_seed(1)).to(cuda) if with_bg else None
    nat = m.native()
    off = ray_offsets(g["ri"], nr)
    sigma, _, rgb, _ = nat.forward(nat.samples_rays(g["o"], g["d"], g["frames"], g["t"], g["ri"]), 1, False, False)
    col = torch.empty(nr, 3, device=cuda)
    ws = torch.empty(nr, device=cuda)
    L.check(L.lib().hrf_composite_forward(sigma.data_ptr(), rgb.data_ptr(), g["t"].data_ptr(), off.data_ptr(), nr, 4e-4, L.ptr(bg),
                                          col.data_ptr(), ws.data_ptr(), None, L.stream()))
    fcol, fws = render_fused(m, g["o"], g["d"], g["frames"], g["t"], g["ri"], nr, bg)
    torch.testing.assert_close(fcol, col, rtol=0, atol=2e-6)
    torch.testing.assert_close(fws.view(-1), ws, rtol=0, atol=2e-6)
    assert (fws.view(-1)[counts.to(cuda) == 0] == 0).all()
    # live count on the device below the capacity: the trailing samples must not be seen
    cut = int(off[100].item())
    count = torch.tensor([cut], dtype=torch.int64, device=cuda)
    off_cut = off.clone()
    off_cut[101:] = cut
    fcol2, fws2 = render_fused(m, g["o"], g["d"], g["frames"], g["t"], g["ri"], nr, bg, count_dev=count, ray_offsets_dev=off_cut)
    torch.testing.assert_close(fcol2[:100], fcol[:100], rtol=0, atol=0)
    assert (fws2.view(-1)[100:] == 0).all()


def test_prune_then_render_reuses_the_features_of_the_density_pass(cuda):
    """prune_samples stores the composed features of the candidates on the batch; render() of that unchanged batch runs
    the MLPs on them (no second encode): bit-identical to a render() that encodes again, with and without gradients; an
    edited batch silently falls back to encoding."""
    from helpers import input_batch_of, make_pair
    from humanrf_b200.volume_rendering import prune_samples, render

    _, m, frames = make_pair((6, 6), table_std=2.0)
    b = synthetic_rays(200, 300, frames, seed=17, ragged=True)
    ib = input_batch_of(b, cuda)
    prune_samples(ib, m, is_training=False)
    from humanrf_b200.volume_rendering import _reusable_features
    assert _reusable_features(ib, m) is not None and set(vars(ib)) == set(vars(input_batch_of(b, cuda)))   # nothing stuck on the batch
    bg = torch.rand(200, 3, device=cuda)
    with torch.no_grad():
        a = render(ib, m, bg, is_training=False)
    ib2 = input_batch_of(b, cuda)
    ib2.sample_distances, ib2.ray_indices = ib.sample_distances.clone(), ib.ray_indices.clone()   # same survivors, no stash
    with torch.no_grad():
        c = render(ib2, m, bg, is_training=False)
    torch.testing.assert_close(a.color, c.color, rtol=0, atol=0)
    torch.testing.assert_close(a.weights_sum, c.weights_sum, rtol=0, atol=0)
    # with gradients: same forward values, gradients equal up to the scatter's atomic ordering
    grads = []
    for batch in (ib, ib2):
        for p in m.hot_parameters():
            p.grad = None
        out = render(batch, m, bg, is_training=True)
        (out.color.square().sum() + out.weights_sum.sum()).backward()
        grads.append([p.grad.clone() for p in m.hot_parameters()])
        torch.testing.assert_close(out.color.detach(), a.color, rtol=0, atol=2e-6)
    for x, y in zip(*grads):      # (the re-gathering scatter blends fp32 table values; the other one reads bf16-rounded per-grid features)
        assert (x - y).norm() <= 5e-3 * y.norm() + 1e-12
