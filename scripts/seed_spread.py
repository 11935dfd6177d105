"""Per-rank work spread of the DP bench: bench.py gives rank r the ray batch of seed 123 + r and reports the max over
ranks.  This runs the same train step on ONE GPU for the batches of ranks 0..7 and prints survivors and phase times, to
tell batch heterogeneity from exchange cost in the N > 1 lines."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from humanrf_b200.dataset.input_batch import InputBatch  # noqa: E402
from humanrf_b200.training import FusedTrainer  # noqa: E402
from humanrf_b200.volume_rendering import render  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
model, frames, _ = bench.build_workload(dev, seed=123)
trainer = FusedTrainer(model, lr=1e-6, reuse=os.environ.get("HRF_TRAIN_REUSE", "feat+grid"))
trainer.profile = True
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
from humanrf_b200.synthetic import synthetic_rays  # noqa: E402

for seed in range(123, 131):
    b = synthetic_rays(bench.RAYS, bench.SPR, frames, seed=seed)
    g = {k: v.to(dev).contiguous() for k, v in b.items() if k in ("o", "d", "frames", "t", "ri")}
    with torch.no_grad():
        out = render(InputBatch(ray_origins=g["o"], ray_directions=g["d"], frame_numbers=g["frames"].view(-1, 1),
                                sample_distances=g["t"].view(-1, 1), ray_indices=g["ri"]), model, None, is_training=False)
        w = out.weights_sum.clamp(min=1e-6)
        g["rgba"] = torch.cat((out.color / w, out.weights_sum), dim=1).clamp(0, 1)
    ms, ph, kept = [], {}, 0
    for i in range(11):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        trainer.step(g["o"], g["d"], g["frames"], g["t"], g["ri"], g["rgba"], bench.RAYS)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ms.append(e0.elapsed_time(e1))
            marks = trainer.last["marks"]
            for (_, a), (name, c) in zip(marks[:-1], marks[1:]):
                ph[name] = ph.get(name, 0.0) + a.elapsed_time(c) / 8
            kept = int(trainer.last["samples"])
    ms.sort()
    print(f"seed {seed} candidates {g['t'].shape[0]} survivors {kept} step_ms {ms[len(ms) // 2]:.4f} "
          + " ".join(f"{k}={v:.3f}" for k, v in sorted(ph.items())), flush=True)
