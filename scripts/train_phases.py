"""Per-phase device time of FusedTrainer.step on the bench workload (diagnostic)."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import bench
from humanrf_b200.training import FusedTrainer
dev = torch.device("cuda:0")
model, frames, b = bench.build_workload(dev, 123)
g = {k: v.to(dev).contiguous() for k, v in b.items() if k in ("o", "d", "frames", "t", "ri", "rgba")}
tr = FusedTrainer(model)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for do_flush in (False, True):
    for i in range(8):
        if do_flush: flush.zero_()
        tr.profile = i >= 3
        torch.cuda.synchronize(); t0 = time.perf_counter()
        tr.step(g["o"], g["d"], g["frames"], g["t"], g["ri"], g["rgba"], bench.RAYS)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
        if i >= 5:
            print("flush" if do_flush else "warm ", f"wall {dt:6.2f} ms samples {tr.last['samples']}", {k: round(v, 2) for k, v in tr.last["phases_ms"].items()})
