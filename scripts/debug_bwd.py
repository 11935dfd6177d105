import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
from helpers import make_pair, positions_of, synthetic_rays
from humanrf_b200 import _lib as L
from oracle import field as F

cuda = torch.device("cuda:0")
om, m, frames = make_pair((6,))
for p in om.parameters():
    p.requires_grad_(True)
q = om._q
for ragged in (False, True):
    b = synthetic_rays(80, 32, frames, ragged=ragged, seed=11)
    pos, dirs, fr = positions_of(b), b["d"][b["ri"]], b["frames"][b["ri"]]
    n = pos.shape[0]
    g = torch.Generator().manual_seed(5)
    for p in om.parameters():
        p.grad = None
    d_sigma = torch.randn(n, generator=g) * 1e-2
    d_rgb = torch.randn(n, 3, generator=g)
    feats = om.features(pos, fr); feats.retain_grad()
    w1, w2 = [q(w) for w in om.w_sigma]
    h = q(torch.relu(feats @ w1.t())); o = h @ w2.t(); o.retain_grad()
    sigma = F.truncated_exp(o[:, 0]) * 100.0
    inp = torch.cat((q(F.sh4(dirs)), q(o[:, 1:]), torch.ones(n, 1)), 1); inp.retain_grad()
    c1, c2, c3 = [q(w) for w in om.w_color]
    h1 = q(torch.relu(inp @ c1.t())); h2 = q(torch.relu(h1 @ c2.t())); h2.retain_grad()
    o3 = h2 @ c3.t(); o3.retain_grad()
    rgb = torch.sigmoid(o3[:, :3])
    ((sigma * d_sigma).sum() + (rgb * d_rgb).sum()).backward()
    nat = m.native()
    s = nat.samples_query(pos.to(cuda).contiguous(), dirs.to(cuda).contiguous(), fr.to(cuda).contiguous())
    _, _, _, feat = nat.forward(s, 1, want_geo=False, want_feat=True)
    params = m.hot_parameters()
    grads = [torch.zeros_like(p) for p in params]
    dbg = torch.zeros(n, 64, device=cuda)
    L.lib().hrf_debug_set_buffer(dbg.data_ptr())
    nat.backward(s, d_sigma.to(cuda), d_rgb.to(cuda).contiguous(), feat, grads)
    torch.cuda.synchronize()
    L.lib().hrf_debug_set_buffer(None)
    D = dbg.cpu()
    def cmp(name, a, r):
        a, r = a.double(), r.double()
        print(f"   {name:10s} max|ref| {r.abs().max():.3e} max|diff| {(a-r).abs().max():.3e} relnorm {((a-r).norm()/r.norm().clamp_min(1e-30)):.3e}")
    print("ragged", ragged, "n", n)
    cmp("h0", D[:, 0], o[:, 0].detach())
    cmp("dh0", D[:, 1], o.grad[:, 0])
    cmp("dgeo", D[:, 2:17], o.grad[:, 1:])
    cmp("o3", D[:, 18:21], o3[:, :3].detach())
    cmp("d3", D[:, 21:24], o3.grad[:, :3])
    cmp("dH2[:8]", D[:, 24:32], h2.grad[:, :8])
    cmp("feat[:8]", D[:, 32:40], feats.detach()[:, :8])
    cmp("L1s[:8]", D[:, 40:48], (feats @ w1.t()).detach()[:, :8])
    print("    W1s row0[:8] smem", D[0, 48:56].tolist(), " ref", w1[0, :8].tolist())
    cmp("L2s[:8]", D[:, 56:64], o.detach()[:, :8])
    refs = [om.segments[0].grids[k].grad.reshape(-1) for k in range(4)] + [om.segments[0].vectors.grad.reshape(-1)]
    refs += [torch.cat([w.grad.reshape(-1) for w in om.w_sigma]), torch.cat([w.grad.reshape(-1) for w in om.w_color])]
    names = ["xyz", "xyt", "yzt", "xzt", "vectors", "sigma_net", "color_net"]
    for nm, got, ref in zip(names, grads, refs):
        got = got.cpu().reshape(-1).double(); ref = ref.double()
        cos = float((got @ ref) / (got.norm() * ref.norm() + 1e-30))
        print(f"  {nm:10s} |got| {got.norm():.4e} |ref| {ref.norm():.4e} ratio {got.norm()/ref.norm():.4f} cos {cos:.5f} nnz got {(got!=0).sum().item()} ref {(ref!=0).sum().item()}")
    for li, (a, bnd) in enumerate([(0, 2048), (2048, 3072)]):
        gg, rr = grads[5].cpu()[a:bnd].double(), refs[5][a:bnd].double()
        print(f"   sigma layer{li} ratio {gg.norm()/rr.norm():.4f} cos {float((gg@rr)/(gg.norm()*rr.norm()+1e-30)):.5f}")
    for li, (a, bnd) in enumerate([(0, 2048), (2048, 6144), (6144, 7168)]):
        gg, rr = grads[6].cpu()[a:bnd].double(), refs[6][a:bnd].double()
        print(f"   color layer{li} ratio {gg.norm()/rr.norm():.4f} cos {float((gg@rr)/(gg.norm()*rr.norm()+1e-30)):.5f}")
