"""TEST INFRASTRUCTURE / BASELINE ONLY.  Times the CPU oracle (the port of the reference's path) on the host cores.
  --mode render: encode -> MLPs -> composite of a ray batch (humanrf.py:188-208, volume_rendering.py:123-145)
  --mode train : prune pass (density of every candidate + visibility, volume_rendering.py:42-84), forward of the survivors,
                 loss (trainer.py:205-215), backward by torch autograd (table gradients as sparse tensors), and ONE dense
                 Adam update of all parameters (run.py:101-104) in the parent with every host thread
The sample batch is split by rays over a fork()ed process pool (one single-threaded worker per core; the model is built
once and shared copy-on-write).  The number of rays per worker is chosen from a calibration step so that the whole
run fits `--budget-s` (never fewer than 16 rays per worker unless a 16-ray step takes seconds: below that, process-pool overhead dominates and the figure
is not reproducible).  Prints one JSON line.  Run as a separate process by bench.py so that no CUDA context is forked."""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

_STATE = {}


def _work(chunk):
    from oracle import rendering as orender

    om, b, mode = _STATE["om"], _STATE["b"], _STATE["mode"]
    r0, r1 = chunk
    torch.set_num_threads(1)
    sel = (b["ri"] >= r0) & (b["ri"] < r1)
    ri = b["ri"][sel]
    t = b["t"][sel]
    pos = b["o"][ri] + t.unsqueeze(1) * b["d"][ri]
    if mode == "render":
        with torch.no_grad():
            sig, _, rgb = om.forward(pos, b["d"][ri], b["frames"][ri])
            col, _ = orender.render(t, sig, rgb, ri - r0, r1 - r0, None)
        return float(col.sum())
    with torch.no_grad():                                    # prune_samples
        sig0, _ = om.density(pos, b["frames"][ri])
        keep = orender.prune_mask(sig0, ri)
    ri, t, pos = ri[keep], t[keep], pos[keep]
    for p in om.parameters():
        p.grad = None
    sig, _, rgb = om.forward(pos, b["d"][ri], b["frames"][ri])
    bg = torch.rand(r1 - r0, 3)
    col, ws = orender.render(t, sig, rgb, ri - r0, r1 - r0, bg)
    loss, _ = orender.training_loss(col, ws, b["rgba"][r0:r1], bg)
    loss.backward()
    return float(loss.detach())


def _step(pool, chunks, adam):
    out = pool.map(_work, chunks)
    if adam is not None:
        adam()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="render", choices=["render", "train"])
    ap.add_argument("--rays-per-worker", type=int, default=0, help="0 = calibrate against --budget-s")
    ap.add_argument("--budget-s", type=float, default=20.0)
    ap.add_argument("--samples-per-ray", type=int, default=512)
    ap.add_argument("--workers", type=int, default=os.cpu_count())
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--segments", type=int, nargs="+", default=[50])
    a = ap.parse_args()
    from humanrf_b200.synthetic import synthetic_rays
    from oracle import field as ofield

    frames = tuple(range(15, 15 + sum(a.segments)))
    train = a.mode == "train"
    om = ofield.make_model(tuple(a.segments), frames, seed=123, table_init="trained", bf16=False, requires_grad=train)
    om.sparse_grad = train
    _STATE["om"], _STATE["mode"] = om, a.mode
    adam = None
    if train:     # the optimiser's share of a step: dense Adam over every parameter, all host threads (parent process)
        params = [p.detach().clone().requires_grad_(True) for p in om.parameters()]
        for p in params:
            p.grad = torch.zeros_like(p)
        opt = torch.optim.Adam(params, lr=1e-2, betas=(0.9, 0.99), eps=1e-15)

        def adam():
            torch.set_num_threads(a.workers)
            opt.step()
            torch.set_num_threads(1)

    torch.set_num_threads(1)
    ctx = mp.get_context("fork")

    def run(rays_per_worker, steps, warmup):
        rays = rays_per_worker * a.workers
        _STATE["b"] = synthetic_rays(rays, a.samples_per_ray, frames, seed=7)
        chunks = [(i * rays_per_worker, (i + 1) * rays_per_worker) for i in range(a.workers)]
        with ctx.Pool(a.workers) as pool:
            for _ in range(warmup):
                _step(pool, chunks, adam)
            t0 = time.perf_counter()
            for _ in range(steps):
                _step(pool, chunks, adam)
            return rays, (time.perf_counter() - t0) / max(steps, 1)

    rpw = a.rays_per_worker
    if rpw <= 0:
        _, sec = run(16, 1, 1)                                # calibration (page-in + one timed step at 16 rays per worker)
        per_step = a.budget_s / max(a.steps + a.warmup, 1)
        # Floor: 16 rays per worker, unless a step of that size takes long enough (train: ~10 s on 128 cores) that a smaller
        # one still dwarfs the process-pool overhead: then as few rays as keep a step above ~1.5 s, so that a run with many
        # steps (--steps 20 --warmup 5) still ends within the budget instead of 25 x 10 s.
        floor = int(min(16, max(1, -(-16 * 1.5 // max(sec, 1e-3)))))
        rpw = int(min(64, max(floor, 16 * per_step / max(sec, 1e-3))))
    rays, sec = run(rpw, a.steps, max(a.warmup, 1))
    what = ("prune pass + forward + loss + autograd backward per worker, one dense Adam over all parameters with all threads"
            if train else "encode + MLPs + composite per worker")
    print(json.dumps({"rays_per_s": rays / sec, "cores": a.workers, "seconds_per_step": sec, "rays_per_step": rays,
                      "rays_per_worker": rpw,
                      "sample": f"{rays} rays x {a.samples_per_ray} samples of the same workload per step ({rpw} rays per worker), fp32, "
                                f"{a.workers} single-threaded workers, {what}, mean of {a.steps} steps"}))


if __name__ == "__main__":
    main()
