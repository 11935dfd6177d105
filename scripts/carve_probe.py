"""Times hrf_occupancy_from_masks at G=256 / 24 cameras for several HRF_CARVE_CHUNKS settings (one process each)."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
if len(sys.argv) > 1:
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    import torch
    from humanrf_b200.synthetic_scene import carve_scene

    from humanrf_b200.toolbox import occupancy_grid_generation_native as ours

    sc = carve_scene(num_cameras=24, width=512, height=384, seed=5)
    cuda = torch.device("cuda", 0)
    args = (torch.from_numpy(sc["masks"]).to(cuda), torch.from_numpy(sc["projection_matrices"]).to(cuda),
            torch.from_numpy(sc["landscape"]).to(cuda), 20, 256, 512, 384)
    for _ in range(3):
        ours.generate_from_masks(*args)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ours.generate_from_masks(*args)
    e1.record(); torch.cuda.synchronize()
    print(f"HRF_CARVE_CHUNKS={os.environ.get('HRF_CARVE_CHUNKS')}: {e0.elapsed_time(e1) / 20:.3f} ms")
else:
    for c in ("1", "4", "16", "64", "100000"):
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, HRF_CARVE_CHUNKS=c))
