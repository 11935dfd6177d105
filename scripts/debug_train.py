import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
from humanrf_b200.synthetic import make_model, synthetic_rays, input_batch_of
from humanrf_b200.volume_rendering import prune_samples, render, ray_offsets
from humanrf_b200.training import FusedTrainer
cuda = torch.device("cuda:0")
def P(*a):
    torch.cuda.synchronize(); print(*a, flush=True)
for segs in ((6,), (6, 6)):
    model, frames = make_model(segs, table_std=0.5, device=cuda)
    for ragged in (False, True):
        b = synthetic_rays(256, 128, frames, seed=5, ragged=ragged)
        ib = input_batch_of(b, cuda)
        P("segs", segs, "ragged", ragged, "n", ib.num_samples)
        nat = model.native()
        o, d, fr, t, ri = ib.ray_origins, ib.ray_directions, ib.frame_numbers.view(-1).contiguous(), ib.sample_distances.view(-1).contiguous(), ib.ray_indices
        s = nat.samples_rays(o, d, fr, t, ri)
        sig, _, _, _ = nat.forward(s, 0, False, False); P(" fwd mode0 ok", float(sig.mean()))
        sig, _, rgb, feat = nat.forward(s, 1, False, True); P(" fwd mode1 ok", float(rgb.mean()))
        prune_samples(ib, model, is_training=True); P(" prune ok", ib.num_samples)
        out = render(ib, model, torch.rand(256, 3, device=cuda), True); P(" render ok")
        out.color.sum().backward(); P(" backward ok")
        g = {k: v.to(cuda).contiguous() for k, v in b.items() if k in ("o", "d", "frames", "t", "ri", "rgba")}
        tr = FusedTrainer(model, prune=True)
        for i in range(2):
            l = tr.step(g["o"], g["d"], g["frames"], g["t"], g["ri"], g["rgba"], 256, return_loss=True); P(" trainer step", i, l, tr.last["samples"])
print("done")
