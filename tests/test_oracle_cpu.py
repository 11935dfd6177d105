"""CPU suite: properties that pin the oracle's own restatement (no reference execution possible for the
tcnn / nerfacc / texture parts: parity unpinned, see oracle/__init__.py)."""
import numpy as np
import pytest
import torch

from helpers import positions_of, synthetic_rays
from oracle import field as F
from oracle import hashgrid, rendering as R, sampler as S


def test_hashgrid_vertex_interpolation_and_layout():
    l2 = 15
    sc, rs, of, sz, hs, tot = hashgrid.level_table(l2)
    g = torch.Generator().manual_seed(0)
    table = torch.randn(tot, 2, generator=g)
    # a sample exactly on a level-0 vertex returns that entry (dense index x + y*32 + z*32^2)
    x = (torch.tensor([[3.0, 7.0, 11.0]]) - 0.5) / float(sc[0])
    enc = hashgrid.encode(table, x, l2)
    idx = 3 + 7 * 32 + 11 * 32 * 32
    torch.testing.assert_close(enc[0, :2], table[idx], rtol=1e-5, atol=1e-5)
    # hashed level: indices stay inside the level and weights sum to one
    xs = torch.rand(1000, 3, generator=g).numpy().astype(np.float32)
    for l in (0, 2, 3, 9, 15):
        idx, w = hashgrid.corner_indices_and_weights(xs, sc[l], int(rs[l]), int(sz[l]), bool(hs[l]))
        assert idx.min() >= 0 and idx.max() < sz[l]
        np.testing.assert_allclose(w.sum(1), 1.0, atol=1e-5)
    assert enc.shape == (1, 32)


def test_compose_pairing_and_vector_lerp():
    # tensor_composition.cu:49-52 : xyz<->t, xyt<->z, yzt<->x, xzt<->y
    vec = torch.zeros(4, 2048, 32)
    vec[0] += 2.0; vec[1] += 3.0; vec[2] += 5.0; vec[3] += 7.0
    one = torch.ones(4, 32)
    c = torch.rand(4, 4)
    out = F.compose(one * 1, one * 10, one * 100, one * 1000, vec, c)
    torch.testing.assert_close(out, torch.full((4, 32), 7.0 + 50.0 + 200.0 + 3000.0))
    # lerp: coordinate of texel centre i returns row i exactly; clamps at both ends
    vec = torch.arange(2048.0).view(1, 2048, 1).repeat(4, 1, 32)
    coords = torch.tensor([[(5 + 0.5) / 2048, 0.0, 1.0, (9 + 1.0) / 2048]])
    sv = F.lerp_vectors(vec, coords)
    assert sv[0, 0, 0] == 5.0 and sv[1, 0, 0] == 0.0 and sv[2, 0, 0] == 2047.0 and abs(sv[3, 0, 0] - 9.5) < 1e-4


def test_frame_luts_follow_reference_rules():
    f2s, f2t = F.frame_luts(tuple(range(15, 65)), (25, 50))     # last segment clipped to the 50 frames (humanrf.py:80)
    assert f2s[14] == -1 and f2s[15] == 0 and f2s[39] == 0 and f2s[40] == 1 and f2s[64] == 1
    assert f2t[15] == 0.0 and abs(f2t[16] - 1 / 25) < 1e-7 and abs(f2t[41] - 1 / 25) < 1e-7   # denominators = actual frames


def test_rendering_identities():
    b = synthetic_rays(50, 40, tuple(range(15, 21)), ragged=True)
    g = torch.Generator().manual_seed(0)
    sigma = torch.rand(b["t"].shape[0], generator=g) * 300
    w = R.weights_from_density(b["t"], sigma, b["ri"])
    ws = R.accumulate(w, b["ri"], None, 50)
    assert (ws <= 1 + 1e-5).all() and (w >= 0).all()
    # sum of weights + final transmittance == 1 per ray
    dt = (b["t"] + 4e-4) - b["t"]
    tot = torch.zeros(50).index_add(0, b["ri"], sigma * dt)
    torch.testing.assert_close(ws[:, 0] + torch.exp(-tot), torch.ones(50), rtol=1e-4, atol=1e-4)
    keep = R.prune_mask(sigma, b["ri"])
    alphas = 1 - torch.exp(-sigma * 4e-4)
    assert not keep[alphas < 1e-4].any()
    col, _ = R.render(b["t"], sigma, torch.ones(sigma.shape[0], 3), b["ri"], 50, torch.zeros(50, 3))
    torch.testing.assert_close(col, ws.expand(-1, 3), rtol=1e-5, atol=1e-6)


def test_texture_emulation_basics():
    G = 8
    grid = np.zeros((G, G, G), np.uint8)
    grid[2, 3, 4] = 255                      # [z][y][x]
    c = lambda i: (i + 0.5) / G
    p = np.array([[c(4), c(3), c(2)], [c(5), c(3), c(2)], [c(4.99), c(3), c(2)], [c(5.0) - 0.4 / (256 * G), c(3), c(2)],
                  [c(3.01), c(3), c(2)], [c(3.0), c(3), c(2)], [c(4), c(3), c(4)]], np.float32)
    np.testing.assert_array_equal(S.tex_occupied(grid, p), [True, False, True, False, True, False, False])


def test_sampler_oracle_against_dense_brute_force():
    """AABB slab result equals a float64 analytic intersection to ~1e-6; sample counts equal the definition."""
    rng = np.random.default_rng(0)
    o = rng.normal(size=(200, 3)); o = (2 * o / np.linalg.norm(o, axis=1, keepdims=True)).astype(np.float32)
    d = (rng.uniform(-0.3, 0.3, (200, 3)) - o); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    aabb = np.array([[-0.5, -0.5, -0.5], [0.5, 0.5, 0.5]], np.float32)
    tmin, tmax = S.aabb_minmax(o, d, aabb)
    t0 = (aabb[0] - o.astype(np.float64)) / d; t1 = (aabb[1] - o.astype(np.float64)) / d
    np.testing.assert_allclose(tmin, np.minimum(t0, t1).max(1), rtol=1e-5)
    np.testing.assert_allclose(tmax, np.maximum(t0, t1).min(1), rtol=1e-5)
    assert (tmin < tmax).all()


def test_oracle_bf16_mode_close_to_fp32_mode():
    om32 = F.make_model((6,), seed=1, bf16=False, table_std=6.0)
    om16 = F.make_model((6,), seed=1, bf16=True, table_std=6.0)
    b = synthetic_rays(20, 10, tuple(range(15, 21)))
    pos, dirs, fr = positions_of(b), b["d"][b["ri"]], b["frames"][b["ri"]]
    with torch.no_grad():
        s32, _, c32 = om32.forward(pos, dirs, fr)
        s16, _, c16 = om16.forward(pos, dirs, fr)
    assert ((s32 - s16).abs() / s32).max() < 0.1 and (c32 - c16).abs().max() < 2e-2
    assert s32.std() / s32.mean() > 0.2, "synthetic parameters must give non-degenerate densities"


def test_config0_cpu_plumbing_sampler_to_input_batch_to_merge():
    """BASELINE configs[0]: oracle sampler -> InputBatch -> merge_input_batches on the CPU, against frozen golden counts."""
    from pathlib import Path

    from humanrf_b200.dataset.input_batch import InputBatch
    from humanrf_b200.input import merge_input_batches
    from scene import make_scene

    G0 = np.load(Path(__file__).resolve().parent / "golden" / "sampler_config0.npz")
    sc = make_scene(num_images=1, width=96, height=72, G=64, portrait_every=0, seed=3)
    batches = []
    for name, occ, step in (("occ", True, 4e-4), ("aabb", False, 4e-3)):
        o, d, rgba, fn, cn, mm, mask, t, rel = S.get_data(sc["rgba"], sc["light_mask"], sc["frame_numbers"], sc["camera_numbers"],
                                                          sc["grids"], sc["landscape"], G0["idx"], sc["inverse_krs"],
                                                          sc["camera_origins"], sc["aabb"], sc["G"], 96, 72, step, False, occupancy=occ)
        np.testing.assert_array_equal(mask, G0[f"{name}_mask"])
        np.testing.assert_array_equal(mm, G0[f"{name}_minmax"])
        np.testing.assert_array_equal(d, G0[f"{name}_dirs"])
        np.testing.assert_array_equal(np.bincount(rel, minlength=mm.shape[0]), G0[f"{name}_counts"])
        np.testing.assert_array_equal(t[:64], G0[f"{name}_t_head"])
        assert (np.diff(rel) >= 0).all() and (mm[:, 0] < mm[:, 1]).all()
        tt = torch.from_numpy
        batches.append(InputBatch(ray_origins=tt(o), ray_directions=tt(d), minmaxes=tt(mm), rgba=tt(rgba),
                                  ray_masks=tt(mask).view(-1, 1), frame_numbers=tt(fn).view(-1, 1),
                                  unique_frame_numbers=torch.unique(tt(fn)).view(-1, 1), camera_numbers=tt(cn).view(-1, 1),
                                  sample_distances=tt(t).view(-1, 1), ray_indices=tt(rel).long(), width=96, height=72))
    merged = merge_input_batches(batches, max_num_samples=int(0.8 * sum(b.num_samples for b in batches)))
    assert merged.num_rays <= sum(b.num_rays for b in batches) and merged.num_samples <= 0.8 * sum(b.num_samples for b in batches)
    assert merged.ray_indices.max() < merged.num_rays and (merged.ray_indices[1:] >= merged.ray_indices[:-1]).all()
    # input.py:41 keeps `cumsum < cutoff`, i.e. one True fewer than surviving rays (reference quirk, mirrored)
    assert int(merged.ray_masks.sum()) in (merged.num_rays - 1, merged.num_rays)
    # AABB rays are a superset of the occupancy rays and their intervals contain the occupancy intervals
    assert (G0["aabb_mask"] | ~G0["occ_mask"]).all()
