"""Where does the end-to-end render step spend its time?  (H2D bandwidth, host enqueue cost of render(), stream depth.)
Run on the GPU box:  python scripts/e2e_probe.py"""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import RAYS, SEGMENTS, SPR  # noqa: E402

from humanrf_b200.dataset.input_batch import InputBatch  # noqa: E402
from humanrf_b200.synthetic import make_model, synthetic_rays  # noqa: E402
from humanrf_b200.volume_rendering import render  # noqa: E402

dev = torch.device("cuda", 0)
model, frames = make_model(SEGMENTS, seed=123, device=dev)
b = synthetic_rays(RAYS, SPR, frames, seed=123)
keys = ("o", "d", "frames", "t", "ri")
host = {k: b[k].contiguous().pin_memory() for k in keys}
bg = torch.rand(RAYS, 3, device=dev)
nbytes = sum(v.numel() * v.element_size() for v in host.values())


def sync_time(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


up = lambda: {k: v.to(dev, non_blocking=True) for k, v in host.items()}  # noqa: E731
ms = sync_time(up)
print(f"H2D {nbytes / 1e6:.1f} MB: {ms:.3f} ms = {nbytes / ms / 1e6:.1f} GB/s")
g = up()
torch.cuda.synchronize()


def call(bb):
    ib = InputBatch(ray_origins=bb["o"], ray_directions=bb["d"], frame_numbers=bb["frames"].view(-1, 1),
                    sample_distances=bb["t"].view(-1, 1), ray_indices=bb["ri"])
    with torch.no_grad():
        return render(ib, model, bg, is_training=False)


print(f"render() on resident inputs, synced: {sync_time(lambda: call(g)):.3f} ms/step")
# host enqueue cost: short bursts that fit the launch queue
call(g); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(8):
    call(g)
enq = (time.perf_counter() - t0) / 8 * 1e3
torch.cuda.synchronize()
print(f"render() host enqueue: {enq:.3f} ms/step")
t0 = time.perf_counter()
for _ in range(8):
    up()
print(f"upload host enqueue: {(time.perf_counter() - t0) / 8 * 1e3:.3f} ms/step")
torch.cuda.synchronize()
for depth in (1, 2, 3, 4):
    streams = [torch.cuda.Stream(dev) for _ in range(depth)]
    outs = [torch.empty(RAYS, 3).pin_memory() for _ in range(depth)]

    def loop(k):
        for i in range(k):
            with torch.cuda.stream(streams[i % depth]):
                bb = up()
                outs[i % depth].copy_(call(bb).color, non_blocking=True)
        torch.cuda.synchronize()

    loop(4)
    t0 = time.perf_counter()
    loop(20)
    dt = (time.perf_counter() - t0) / 20 * 1e3
    print(f"depth {depth}: {dt:.3f} ms/step = {RAYS / dt / 1e3:.2f} M rays/s")
