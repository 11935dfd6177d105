"""Per-phase device time of one full-image render (bench.py --mode image workload): sampler, density pass, pruning, fused
MLP + compositing, placement.  Diagnostic for DESIGN.md section 4."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from humanrf_b200 import _lib as L
from humanrf_b200.dataset import ray_sampler_native as rs
from humanrf_b200.dataset.occupancy_grid_native import OccupanyGrid
from humanrf_b200.synthetic import make_model
from humanrf_b200.synthetic_scene import make_scene
from humanrf_b200.volume_rendering import ray_offsets, render_fused

dev = torch.device("cuda:0")
W, H, G = 1028, 752, 256
model, frames = make_model((50,), seed=123, device=dev)
sc = make_scene(num_images=1, width=W, height=H, G=G, portrait_every=0)
og = OccupanyGrid(G, 1)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
cam = dict(frame_numbers=t(sc["frame_numbers"]), camera_numbers=t(sc["camera_numbers"]),
           grid_handles=torch.tensor([og.add_grid(t(sc["grids"][0]))], dtype=torch.int64, device=dev), landscape=t(sc["landscape"]),
           inverse_krs=t(sc["inverse_krs"]), camera_origins=t(sc["camera_origins"]), aabb=t(sc["aabb"]))
nat, lib, step = model.native(), L.lib(), 4e-4
empty_rgba, empty_mask = torch.zeros((0, 4), dtype=torch.uint8, device=dev), torch.zeros(0, dtype=torch.bool, device=dev)
acc = {}


def mark(name, ev_prev):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return name, ev_prev, e


for it in range(4):
    evs = []
    e0 = torch.cuda.Event(enable_timing=True); e0.record()
    prev = e0
    stats = dict(candidates=0, rays=0)
    for s in range(0, W * H, 262144):
        e = min(s + 262144, W * H)
        idx = torch.arange(s, e, dtype=torch.int64, device=dev)
        (o, d, _, fn, cn, mm, mask, dist_, rel) = rs.get_samples_occupancy_minmax(
            empty_rgba, empty_mask, cam["frame_numbers"], cam["camera_numbers"], cam["grid_handles"], cam["landscape"], idx,
            cam["inverse_krs"], cam["camera_origins"], cam["aabb"], G, W, H, step, False)
        n_, prev_, ev = mark("sampler", prev); evs.append((n_, prev_, ev)); prev = ev
        nr, n = o.shape[0], dist_.shape[0]
        stats["candidates"] += n; stats["rays"] += nr
        ri = rel.long()
        off0 = ray_offsets(ri, nr)
        samples = nat.samples_rays(o, d, fn, dist_, ri)
        sigma, saved = nat.density_early_stop(samples, off0, nr, step, save="feat")
        n_, prev_, ev = mark("density pass", prev); evs.append((n_, prev_, ev)); prev = ev
        keep = torch.empty(n, dtype=torch.uint8, device=dev); kept_off = torch.empty(nr + 1, dtype=torch.int32, device=dev)
        t2 = torch.empty(n, device=dev); ri2 = torch.empty(n, dtype=torch.int64, device=dev); src = torch.empty(n, dtype=torch.int32, device=dev)
        count = torch.zeros(1, dtype=torch.int64, device=dev)
        L.check(lib.hrf_prune(sigma.data_ptr(), dist_.data_ptr(), ri.data_ptr(), off0.data_ptr(), nr, step, 1e-4, 1e-4, keep.data_ptr(),
                              kept_off.data_ptr(), t2.data_ptr(), ri2.data_ptr(), src.data_ptr(), count.data_ptr(), L.stream()))
        n_, prev_, ev = mark("prune", prev); evs.append((n_, prev_, ev)); prev = ev
        color, _ = render_fused(model, o, d, fn, t2, ri2, nr, 0.0, step, reuse=(saved, src, kept_off), count_dev=count)
        n_, prev_, ev = mark("fused MLP+composite", prev); evs.append((n_, prev_, ev)); prev = ev
        stats["survivors"] = stats.get("survivors", 0) + int(count.item())
    torch.cuda.synchronize()
    if it == 3:
        for name, a, b in evs:
            acc[name] = acc.get(name, 0.0) + a.elapsed_time(b)
        print("image phases (ms):", {k: round(v, 3) for k, v in acc.items()}, "total", round(e0.elapsed_time(prev), 3), stats)
