"""Size-independent properties at BASELINE's full bench size (4096 rays x 512 samples, segment_sizes=(50,)), where the
CPU oracle would take minutes: split invariance, transmittance identity, pruning idempotence, gradient linearity."""
import numpy as np
import pytest
import torch

from humanrf_b200 import _lib as L
from humanrf_b200.synthetic import input_batch_of, make_model, synthetic_rays
from humanrf_b200.volume_rendering import prune_samples, ray_offsets, render

pytestmark = pytest.mark.gpu
R, S = 4096, 512


@pytest.fixture(scope="module")
def world():
    model, frames = make_model((50,), device="cuda")
    b = synthetic_rays(R, S, frames, seed=123)
    return model, frames, b


def test_forward_is_invariant_to_how_the_batch_is_split(cuda, world):
    model, frames, b = world
    nat = model.native()
    g = {k: v.to(cuda).contiguous() for k, v in b.items() if k in ("o", "d", "frames", "t", "ri")}
    full = nat.forward(nat.samples_rays(g["o"], g["d"], g["frames"], g["t"], g["ri"]), 1, False, False)
    cut = (R // 3) * S + 77                                   # not a multiple of the 128-sample tile
    parts = [nat.forward(nat.samples_rays(g["o"], g["d"], g["frames"], g["t"][a:e].contiguous(), g["ri"][a:e].contiguous()),
                         1, False, False) for a, e in ((0, cut), (cut, R * S))]
    torch.testing.assert_close(torch.cat([p[0] for p in parts]), full[0], rtol=0, atol=0)
    torch.testing.assert_close(torch.cat([p[2] for p in parts]), full[2], rtol=0, atol=0)
    assert torch.isfinite(full[0]).all() and (full[2] >= 0).all() and (full[2] <= 1).all()


def test_weights_and_transmittance_partition_unity(cuda, world):
    model, frames, b = world
    nat = model.native()
    g = {k: v.to(cuda).contiguous() for k, v in b.items() if k in ("o", "d", "frames", "t", "ri")}
    sigma, _, rgb, _ = nat.forward(nat.samples_rays(g["o"], g["d"], g["frames"], g["t"], g["ri"]), 1, False, False)
    off = ray_offsets(g["ri"], R)
    color, wsum, w = torch.empty(R, 3, device=cuda), torch.empty(R, device=cuda), torch.empty(R * S, device=cuda)
    L.check(L.lib().hrf_composite_forward(sigma.data_ptr(), rgb.data_ptr(), g["t"].data_ptr(), off.data_ptr(), R, 4e-4, None,
                                          color.data_ptr(), wsum.data_ptr(), w.data_ptr(), L.stream()))
    dt = (g["t"] + 4e-4) - g["t"]
    depth = torch.zeros(R, device=cuda, dtype=torch.float64).index_add(0, g["ri"], (sigma * dt).double())
    torch.testing.assert_close(wsum.double() + torch.exp(-depth), torch.ones(R, device=cuda, dtype=torch.float64), rtol=0, atol=2e-5)
    torch.testing.assert_close(w.view(R, S).sum(1), wsum, rtol=1e-5, atol=1e-6)
    assert (color <= wsum[:, None] + 1e-5).all()              # radiance in [0,1]


def test_pruning_is_idempotent_and_order_preserving(cuda, world):
    model, frames, b = world
    ib = input_batch_of(b, cuda)
    prune_samples(ib, model, is_training=False)
    n1, t1, r1 = ib.num_samples, ib.sample_distances.clone(), ib.ray_indices.clone()
    assert 0 < n1 < R * S and (r1[1:] >= r1[:-1]).all()
    prune_samples(ib, model, is_training=False)
    assert ib.num_samples == n1 and torch.equal(ib.sample_distances, t1) and torch.equal(ib.ray_indices, r1)


def test_gradient_is_linear_in_the_upstream_gradient(cuda, world):
    model, frames, b = world
    ib = input_batch_of(b, cuda)
    prune_samples(ib, model, is_training=False)
    bg = torch.rand(R, 3, device=cuda, generator=torch.Generator(device=cuda).manual_seed(1))
    params = model.hot_parameters()
    gs = []
    for scale in (1.0, 3.0):
        for p in params:
            p.grad = None
        out = render(ib, model, bg, is_training=True)
        ((out.color.sum() + out.weights_sum.sum()) * scale).backward()
        gs.append([p.grad.clone() for p in params])
    for a, c in zip(*gs):
        rel = (c - 3.0 * a).norm() / c.norm().clamp_min(1e-30)
        assert rel < 2e-2, rel                                  # bf16 gradient tiles: linear up to rounding
