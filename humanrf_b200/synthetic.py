"""Seeded synthetic workloads: there is no ActorsHQ data and no network, so the bench,
smoke() and the tests all draw rays / parameters from here."""
from __future__ import annotations

import torch

MODEL_KW = dict(density_scale=100.0, n_features_per_level=2, log2_hashmap_size=19, n_levels=16, coarsest_resolution=32,
                finest_resolution=2048, geometry_feature_dim=15, n_neurons=64, n_hidden_layers_density=1,
                n_hidden_layers_color=2, sh_degree=4, camera_embedding_dim=0)


def synthetic_rays(num_rays, samples_per_ray, frames, seed=123, n_distinct_frames=8, ragged=False, step=4e-4):
    """Origins on the sphere |o|=2 aimed at U([-0.35,0.35]^3); t_k = t0 + k*step inside the unit cube;
    frames drawn from `n_distinct_frames` of `frames` (max_num_frames_per_batch, run_args.py:101)."""
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(num_rays, 3, generator=g)
    o = 2.0 * o / o.norm(dim=1, keepdim=True)
    tgt = (torch.rand(num_rays, 3, generator=g) - 0.5) * 0.7
    d = tgt - o
    d = d / d.norm(dim=1, keepdim=True)
    inv = 1.0 / d
    t0, t1 = (-0.5 - o) * inv, (0.5 - o) * inv
    tn = torch.minimum(t0, t1).max(dim=1)[0]
    tf = torch.maximum(t0, t1).min(dim=1)[0]
    if ragged:
        counts = torch.randint(0, samples_per_ray + 1, (num_rays,), generator=g)
        counts[::7] = 0
    else:
        counts = torch.full((num_rays,), samples_per_ray, dtype=torch.int64)
    span = (tf - tn - samples_per_ray * step).clamp(min=0)
    start = tn + torch.rand(num_rays, generator=g) * span
    ri = torch.repeat_interleave(torch.arange(num_rays), counts)
    first = torch.cumsum(counts, 0) - counts
    k = torch.arange(ri.numel()) - first[ri]
    t = (start[ri] + k.float() * step).float()
    pool = torch.tensor(frames)[torch.randperm(len(frames), generator=g)[:min(n_distinct_frames, len(frames))]]
    fr = pool[torch.randint(0, pool.numel(), (num_rays,), generator=g)].to(torch.int32)
    cam = torch.randint(0, 160, (num_rays,), generator=g, dtype=torch.int32)
    rgba = torch.rand(num_rays, 4, generator=g)
    rgba[:, 3] = (rgba[:, 3] > 0.5).float()
    return dict(o=o.float(), d=d.float(), t=t, ri=ri.long(), frames=fr, cams=cam, rgba=rgba, counts=counts)


def make_model(segment_sizes=(50,), first_frame=15, seed=123, table_std=0.05, device="cuda"):
    """HumanRF with 'trained-like' N(0, table_std) tables (tcnn's U(-1e-4,1e-4) init gives degenerate densities)."""
    from .scene_representation.humanrf import HumanRF

    frames = tuple(range(first_frame, first_frame + sum(segment_sizes)))
    torch.manual_seed(seed)
    m = HumanRF(sorted_frame_numbers=frames, segment_sizes=tuple(segment_sizes), **MODEL_KW)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for fg in m.feature_grids:
            for p in fg.grids():
                p.copy_(torch.randn(p.shape, generator=g) * table_std)
    return m.to(device), frames


def input_batch_of(b, device):
    from .dataset.input_batch import InputBatch

    fr = b["frames"].view(-1, 1)
    n = b["o"].shape[0]
    return InputBatch(ray_origins=b["o"].to(device), ray_directions=b["d"].to(device),
                      minmaxes=torch.zeros(n, 2, device=device), rgba=b["rgba"].to(device),
                      ray_masks=torch.ones(n, 1, dtype=torch.bool, device=device), frame_numbers=fr.to(device),
                      unique_frame_numbers=torch.unique(fr).view(-1, 1).to(device),
                      camera_numbers=b["cams"].view(-1, 1).to(device), sample_distances=b["t"].view(-1, 1).to(device),
                      ray_indices=b["ri"].to(device), width=64, height=64)
