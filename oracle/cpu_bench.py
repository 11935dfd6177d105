"""TEST INFRASTRUCTURE / BASELINE ONLY.  Times the CPU oracle (the port of the reference's encode -> MLP ->
composite path) on the host cores: the sample batch is split by rays over a fork()ed process pool (one
single-threaded worker per core; the model is built once and shared copy-on-write).  Prints one JSON line.
Run as a separate process by bench.py so that no CUDA context is forked."""
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

    om, b = _STATE["om"], _STATE["b"]
    r0, r1 = chunk
    torch.set_num_threads(1)
    sel = (b["ri"] >= r0) & (b["ri"] < r1)
    ri = b["ri"][sel]
    t = b["t"][sel]
    pos = b["o"][ri] + t.unsqueeze(1) * b["d"][ri]
    with torch.no_grad():
        sig, _, rgb = om.forward(pos, b["d"][ri], b["frames"][ri])
        col, _ = orender.render(t, sig, rgb, ri - r0, r1 - r0, None)
    return float(col.sum())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays-per-worker", type=int, default=4)
    ap.add_argument("--samples-per-ray", type=int, default=512)
    ap.add_argument("--workers", type=int, default=os.cpu_count())
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--segments", type=int, nargs="+", default=[50])
    a = ap.parse_args()
    from humanrf_b200.synthetic import synthetic_rays
    from oracle import field as ofield

    torch.set_num_threads(1)
    frames = tuple(range(15, 15 + sum(a.segments)))
    _STATE["om"] = ofield.make_model(tuple(a.segments), frames, seed=123, table_init="trained", bf16=False)
    rays = a.rays_per_worker * a.workers
    _STATE["b"] = synthetic_rays(rays, a.samples_per_ray, frames, seed=7)
    chunks = [(i * a.rays_per_worker, (i + 1) * a.rays_per_worker) for i in range(a.workers)]
    ctx = mp.get_context("fork")
    with ctx.Pool(a.workers) as pool:
        for _ in range(max(a.warmup, 1)):
            pool.map(_work, chunks)                      # warm-up (page-in, imports)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            pool.map(_work, chunks)
        total = time.perf_counter() - t0
    print(json.dumps({"rays_per_s": rays * a.steps / total, "cores": a.workers, "seconds_per_step": total / a.steps,
                      "rays_per_step": rays,
                      "sample": f"{rays} rays x {a.samples_per_ray} samples of the same workload per step, fp32, "
                                f"{a.workers} single-threaded workers, mean of {a.steps} steps"}))


if __name__ == "__main__":
    main()
