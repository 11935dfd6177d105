"""Differential tests against the LIVE reference (imported read-only from /root/reference) on randomized inputs, for the
first-party pure-Python pieces of the path.  They complement the frozen goldens of tests/golden/ (which also run on the
GPU box, where the reference does not exist): here every run draws many more cases.  Skipped when the reference is
not mounted."""
import dataclasses
import sys
import types
from pathlib import Path

import numpy as np
import pytest
import torch

REF = Path("/root/reference")
pytestmark = pytest.mark.skipif(not REF.exists(), reason="reference checkout not present (GPU box)")


@pytest.fixture(scope="module")
def ref():
    """The reference's importable modules.  Imported under their own names, so anything of ours that was shadowed is
    restored afterwards."""
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("humanrf", "actorshq")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, str(REF))
    try:
        import actorshq.dataset.input_batch as ib
        import humanrf.input as inp
        import humanrf.scene_representation.query_io as qio
        import humanrf.utils.activation as act
        import humanrf.utils.loss as loss
        # adaptive_temporal_partitioning imports VolumetricDataset only for an annotation (needs cv2 etc.): stub it
        if "actorshq.dataset.volumetric_dataset" not in sys.modules:
            stub = types.ModuleType("actorshq.dataset.volumetric_dataset")
            stub.VolumetricDataset = object
            sys.modules["actorshq.dataset.volumetric_dataset"] = stub
        import humanrf.adaptive_temporal_partitioning as atp
        yield types.SimpleNamespace(ib=ib, inp=inp, qio=qio, act=act, loss=loss, atp=atp)
    finally:
        sys.path.remove(str(REF))
        for k in [k for k in sys.modules if k.split(".")[0] in ("humanrf", "actorshq")]:
            del sys.modules[k]
        sys.modules.update(saved)


def _surface(cls):
    fields = [(f.name, f.default) for f in dataclasses.fields(cls)]
    return fields, sorted(m for m in dir(cls) if not m.startswith("_"))


def test_dataclass_surfaces_match(ref):
    from humanrf_b200.dataset.input_batch import InputBatch
    from humanrf_b200.scene_representation.query_io import QueryInput, QueryOutput

    assert _surface(InputBatch) == _surface(ref.ib.InputBatch)
    assert _surface(QueryInput) == _surface(ref.qio.QueryInput)
    assert _surface(QueryOutput) == _surface(ref.qio.QueryOutput)


FIELDS = ["ray_origins", "ray_directions", "minmaxes", "rgba", "ray_masks", "frame_numbers", "unique_frame_numbers",
          "camera_numbers", "sample_distances", "ray_indices"]


def _batch(cls, rng, num_rays, masked):
    counts = rng.integers(0, 9, num_rays)
    ri = np.repeat(np.arange(num_rays), counts)
    mask = np.ones((num_rays + masked, 1), bool)
    mask[rng.permutation(num_rays + masked)[:masked]] = False
    fr = rng.integers(15, 40, (num_rays, 1)).astype(np.int32)
    t = torch.from_numpy
    return cls(ray_origins=t(rng.normal(size=(num_rays, 3)).astype(np.float32)),
               ray_directions=t(rng.normal(size=(num_rays, 3)).astype(np.float32)),
               minmaxes=t(rng.random((num_rays, 2)).astype(np.float32)), rgba=t(rng.random((num_rays, 4)).astype(np.float32)),
               ray_masks=t(mask), frame_numbers=t(fr), unique_frame_numbers=torch.unique(t(fr)).view(-1, 1),
               camera_numbers=t(rng.integers(0, 160, (num_rays, 1)).astype(np.int32)),
               sample_distances=t(rng.random((int(counts.sum()), 1)).astype(np.float32)), ray_indices=t(ri.astype(np.int64)),
               width=64, height=48)


def test_merge_input_batches_randomized(ref):
    """input.py:10-55 incl. the sample-budget cut-off and its `cumsum < cutoff` behaviour, 60 random configurations."""
    from humanrf_b200.dataset.input_batch import InputBatch
    from humanrf_b200.input import merge_input_batches

    rng = np.random.default_rng(7)
    checked_cut = 0
    for case in range(60):
        nb = int(rng.integers(1, 5))
        specs = [(int(rng.integers(1, 30)), int(rng.integers(0, 5))) for _ in range(nb)]
        seed = int(rng.integers(0, 2 ** 31))
        total = None
        outs = []
        for cls, fn in ((ref.ib.InputBatch, ref.inp.merge_input_batches), (InputBatch, merge_input_batches)):
            r2 = np.random.default_rng(seed)
            batches = [_batch(cls, r2, *s) for s in specs]
            total = sum(b.sample_distances.shape[0] for b in batches)
            budget = [None, max(1, total // 2), max(1, total - 1), total, total + 5, 1][case % 6]
            outs.append(fn(batches, budget))
        a, b = outs
        if budget is not None and budget < total:
            checked_cut += 1
        for f in FIELDS:
            x, y = getattr(a, f), getattr(b, f)
            if f == "unique_frame_numbers":
                x, y = torch.sort(x.reshape(-1))[0], torch.sort(y.reshape(-1))[0]
            assert x.dtype == y.dtype and x.shape == y.shape and torch.equal(x, y), (case, f)
        assert (a.width, a.height) == (b.width, b.height)
    assert checked_cut >= 15


def test_truncated_exp_and_bce_randomized(ref):
    from humanrf_b200.utils.activation import truncated_exp
    from humanrf_b200.utils.loss import bce_loss

    g = torch.Generator().manual_seed(3)
    for scale in (1.0, 9.0, 30.0):
        x = torch.randn(513, generator=g) * scale
        dy = torch.randn(513, generator=g)
        outs = []
        for fn in (ref.act.truncated_exp, truncated_exp):
            xi = x.clone().requires_grad_(True)
            y = fn(xi)
            y.backward(dy)
            outs.append((y.detach(), xi.grad))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    pred = torch.rand(777, 1, generator=g) * 1.6 - 0.3
    target = (torch.rand(777, 1, generator=g) > 0.4).float()
    assert torch.equal(ref.loss.bce_loss(pred, target), bce_loss(pred, target))


def test_segment_size_rules_and_partitioning_randomized(ref):
    """adaptive_temporal_partitioning.py:28-107 against the oracle restatement (the GPU implementation is compared with
    the same reference decisions through tests/golden/partitioning.npz)."""
    from humanrf_b200 import adaptive_temporal_partitioning as ours
    from oracle import occupancy_tools as O

    assert ours.PREDEFINED_SEGMENT_SIZES == ref.atp.PREDEFINED_SEGMENT_SIZES == O.PREDEFINED_SEGMENT_SIZES
    for n in range(1, 260):
        assert ours.get_segment_size(n) == ref.atp.get_segment_size(n) == O.get_segment_size(n)
        assert ours.get_final_segment_size(n) == ref.atp.get_final_segment_size(n) == O.get_final_segment_size(n)

    class DS:
        def __init__(self, grids):
            self.grids = grids

        def get_occupancy_grid(self, frame_number):
            return self.grids[frame_number].copy()       # the reference ORs into the first grid of a cluster in place

    rng = np.random.default_rng(11)
    for case in range(12):
        n = int(rng.integers(3, 140))
        base = rng.random((12, 12, 12)) < 0.2
        grids, cur = [], base.copy()
        for f in range(n):
            if rng.random() < [0.05, 0.3, 0.8][case % 3]:
                cur = cur | (rng.random(cur.shape) < 0.02)    # occasional growth of the occupied set
            if rng.random() < 0.03:
                cur = rng.random(cur.shape) < 0.2             # a jump
            grids.append((cur * 255).astype(np.uint8))
        thr = [1.05, 1.25, 2.0][case % 3]
        import contextlib
        import io

        with contextlib.redirect_stderr(io.StringIO()):      # tqdm bar
            want = ref.atp.compute_adaptive_segment_sizes(DS(grids), list(range(n)), thr)
        assert O.compute_adaptive_segment_sizes(grids, thr) == want, case


_MODEL_CACHE = {}


def _import_reference_model():
    """The reference's HumanRF with `tinycudann` and its compiled extension stubbed out: the stubs only RECORD the
    configs they are constructed with and own one flat `.params` (sized by our layout rule -- sizes of tcnn tensors are
    therefore not evidence, the configs / names / first-party tensors are)."""
    from humanrf_b200.scene_representation.grid_layout import GridLayout, MLP_SIGMA_PARAMS, mlp_color_params

    if _MODEL_CACHE:                       # the module stays bound to the stub classes (and their log) of its first import
        _MODEL_CACHE["calls"].clear()
        sys.modules["tinycudann"] = _MODEL_CACHE["tcnn"]
        sys.modules["humanrf.scene_representation.humanrf"] = _MODEL_CACHE["module"]
        return _MODEL_CACHE["module"], _MODEL_CACHE["calls"]
    calls = []

    class _Flat(torch.nn.Module):
        def __init__(self, n):
            super().__init__()
            self.params = torch.nn.Parameter(torch.zeros(n))

    class Encoding(_Flat):
        def __init__(self, n_input_dims, encoding_config, **kw):
            calls.append(("Encoding", n_input_dims, dict(encoding_config)))
            c = encoding_config
            fin = c["base_resolution"] * c["per_level_scale"] ** (c["n_levels"] - 1)
            super().__init__(GridLayout(c["log2_hashmap_size"], c["n_levels"], c["base_resolution"], int(round(fin))).n_params)

    class Network(_Flat):
        def __init__(self, n_input_dims, n_output_dims, network_config, **kw):
            calls.append(("Network", n_input_dims, n_output_dims, dict(network_config)))
            super().__init__(MLP_SIGMA_PARAMS)

    class NetworkWithInputEncoding(_Flat):
        def __init__(self, n_input_dims, n_output_dims, encoding_config, network_config, **kw):
            calls.append(("NetworkWithInputEncoding", n_input_dims, n_output_dims, dict(encoding_config), dict(network_config)))
            super().__init__(mlp_color_params(n_input_dims - 18))

    tcnn = types.ModuleType("tinycudann")
    tcnn.Encoding, tcnn.Network, tcnn.NetworkWithInputEncoding = Encoding, Network, NetworkWithInputEncoding
    sys.modules["tinycudann"] = tcnn
    sys.modules["humanrf.scene_representation.tensor_composition_native"] = types.ModuleType("tensor_composition_native")
    import humanrf.scene_representation.humanrf as ref_model

    _MODEL_CACHE.update(module=ref_model, calls=calls, tcnn=tcnn)
    return ref_model, calls


@pytest.mark.parametrize("segment_sizes,first,count,cam_emb", [((50,), 15, 50, 0), ((25, 12, 100, 6), 0, 140, 2), ((6, 6), 3, 9, 0)])
def test_model_constructor_against_the_reference(ref, segment_sizes, first, count, cam_emb):
    """humanrf.py:68-156 + decomposition4d.py:73-122 run for real (only tcnn is a recording stub): frame LUTs, per-segment
    hash-map sizes, encoding / network configs, state-dict key names and the shapes of the first-party tensors."""
    from humanrf_b200.scene_representation.grid_layout import GridLayout
    from humanrf_b200.scene_representation.humanrf import HumanRF
    from humanrf_b200.synthetic import MODEL_KW

    try:
        ref_model, calls = _import_reference_model()
        frames = tuple(range(first, first + count))
        kw = {**MODEL_KW, "camera_embedding_dim": cam_emb, "temporal_partitioning": "adaptive", "fixed_segment_size": 6}
        theirs = ref_model.HumanRF(sorted_frame_numbers=frames, segment_sizes=segment_sizes, **kw)
        ours = HumanRF(sorted_frame_numbers=frames, segment_sizes=segment_sizes, **kw)
    finally:
        sys.modules.pop("tinycudann", None)
    # frame -> segment / local-time look-up tables
    assert torch.equal(ours.frame_numbers_to_segment_numbers, theirs.frame_numbers_to_segment_numbers)
    assert torch.equal(ours.frame_numbers_to_normalized_local_frame_numbers, theirs.frame_numbers_to_normalized_local_frame_numbers)
    assert (ours.num_frames, ours.num_segments, ours.total_feature_dim, ours.density_scale) == \
           (theirs.num_frames, theirs.num_segments, theirs.total_feature_dim, theirs.density_scale)
    # what the reference asks tcnn for, segment by segment
    enc = [c for c in calls if c[0] == "Encoding"]
    assert len(enc) == 4 * len(segment_sizes)
    for s, fg in enumerate(ours.feature_grids):
        for c in enc[4 * s:4 * s + 4]:
            cfg = c[2]
            assert c[1] == 3 and cfg["otype"] == "HashGrid" and cfg["n_levels"] == 16 and cfg["n_features_per_level"] == 2
            assert cfg["log2_hashmap_size"] == fg.layout.log2_hashmap_size
            assert cfg["base_resolution"] == 32 and np.float32(cfg["per_level_scale"]) == np.float32(np.exp(np.log(2048 / 32) / 15))
            assert GridLayout(cfg["log2_hashmap_size"]).n_params == fg.layout.n_params
    net = [c for c in calls if c[0] == "Network"][0]
    assert net[1:3] == (32, 16) and net[3] == {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None",
                                              "n_neurons": 64, "n_hidden_layers": 1}
    col = [c for c in calls if c[0] == "NetworkWithInputEncoding"][0]
    assert col[1:3] == (18 + cam_emb, 3) and col[4]["output_activation"] == "Sigmoid" and col[4]["n_hidden_layers"] == 2
    assert col[3]["nested"][0] == {"n_dims_to_encode": 3, "otype": "SphericalHarmonics", "degree": 4}
    # state dict: same keys in the same order, same shapes (first-party tensors: vectors, LUT buffers, embeddings)
    a, b = ours.state_dict(), theirs.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, k
    # optimiser parameter groups (humanrf.py:210-220)
    ga, gb = ours.get_params(1e-2), theirs.get_params(1e-2)
    assert [len(list(g["params"])) for g in ga] == [len(list(g["params"])) for g in gb] and [g["lr"] for g in ga] == [g["lr"] for g in gb]


def test_prune_and_render_glue_against_the_reference(ref):
    """The reference's own prune_samples / render / merge_render_outputs (volume_rendering.py:26-150) executed on the CPU
    with `nerfacc` replaced by the oracle's restatement of its three functions and a closed-form stand-in for the scene
    representation: pins everything the oracle restates AROUND nerfacc (positions, jitter, alpha, t_ends = t + step,
    mask application, background blend, output shapes).  nerfacc's own arithmetic stays unpinned."""
    from helpers import synthetic_rays
    from humanrf_b200.volume_rendering import RenderOutput as OurRenderOutput
    from oracle import rendering as R

    nerfacc = types.ModuleType("nerfacc")
    nerfacc.render_visibility = lambda alphas, ray_indices, early_stop_eps, alpha_thre, n_rays: \
        R.render_visibility(alphas, ray_indices, early_stop_eps, alpha_thre)

    def render_weight_from_density(t_starts, t_ends, sigmas, ray_indices, n_rays):
        sdt = sigmas.reshape(-1) * (t_ends - t_starts).reshape(-1)
        return (torch.exp(-R._exclusive_by_ray(sdt, ray_indices, "sum")) * (1.0 - torch.exp(-sdt))).unsqueeze(-1)

    nerfacc.render_weight_from_density = render_weight_from_density
    nerfacc.accumulate_along_rays = lambda weights, ray_indices, values=None, n_rays=None: R.accumulate(weights, ray_indices, values, n_rays)
    try:
        _import_reference_model()
        sys.modules["nerfacc"] = nerfacc
        import humanrf.volume_rendering as vr
    finally:
        sys.modules.pop("tinycudann", None)
        sys.modules.pop("nerfacc", None)
    QueryOutput = ref.qio.QueryOutput

    class Scene:
        def density(self, q):
            r2 = (q.positions ** 2).sum(1)
            return QueryOutput(density=2500.0 * torch.exp(-60.0 * r2) * (1.0 + 0.25 * (q.frame_numbers.reshape(-1) % 3).float()))

        def __call__(self, q):
            return QueryOutput(density=self.density(q).density, radiance=torch.sigmoid(3.0 * q.positions + q.directions))

    frames = tuple(range(15, 27))
    b = synthetic_rays(64, 40, frames, ragged=True, seed=8)
    o, d, fr, ri = b["o"], b["d"], b["frames"].view(-1, 1), b["ri"]

    def batch(t):
        return ref.ib.InputBatch(ray_origins=o, ray_directions=d, frame_numbers=fr, unique_frame_numbers=torch.unique(fr).view(-1, 1),
                                 camera_numbers=torch.zeros_like(fr), sample_distances=t.clone().view(-1, 1), ray_indices=ri.clone(),
                                 rgba=b["rgba"], width=8, height=8)

    scene = Scene()
    for is_training in (False, True):
        ib = batch(b["t"])
        torch.manual_seed(5)
        vr.prune_samples(ib, scene, is_training=is_training)
        torch.manual_seed(5)
        t = b["t"].view(-1, 1) + (torch.rand_like(b["t"].view(-1, 1)) * R.STEP if is_training else 0)
        pos = o[ri] + t * d[ri]
        sigma = scene.density(types.SimpleNamespace(positions=pos, frame_numbers=fr[ri])).density
        keep = R.prune_mask(sigma, ri)
        assert 0 < int(keep.sum()) < keep.numel()
        assert torch.equal(ib.sample_distances, t[keep]) and torch.equal(ib.ray_indices, ri[keep])
        assert ib.sample_distances.shape == (int(keep.sum()), 1)
        # render the survivors
        bg = torch.rand(64, 3, generator=torch.Generator().manual_seed(1))
        out = vr.render(ib, scene, bg, is_training=is_training)
        tk, rk = t[keep], ri[keep]
        pk = o[rk] + tk * d[rk]
        q = types.SimpleNamespace(positions=pk, directions=d[rk], frame_numbers=fr[rk])
        col, ws = R.render(tk, scene.density(q).density, scene(q).radiance, rk, 64, bg)
        assert torch.equal(out.color, col) and torch.equal(out.weights_sum, ws)
        assert out.color.shape == (64, 3) and out.weights_sum.shape == (64, 1)
        nobg = vr.render(ib, scene, None, is_training=is_training)
        assert torch.equal(nobg.color, R.render(tk, scene.density(q).density, scene(q).radiance, rk, 64, None)[0])
    # merge_render_outputs: same concatenation, same error for a field that is not a tensor
    parts = [vr.RenderOutput(color=torch.rand(n, 3), weights_sum=torch.rand(n, 1)) for n in (3, 0, 5)]
    a = vr.RenderOutput.merge_render_outputs(parts)
    c = OurRenderOutput.merge_render_outputs([OurRenderOutput(color=p.color, weights_sum=p.weights_sum) for p in parts])
    assert torch.equal(a.color, c.color) and torch.equal(a.weights_sum, c.weights_sum)
    for cls in (vr.RenderOutput, OurRenderOutput):
        with pytest.raises(RuntimeError, match="Unknown data type"):
            cls.merge_render_outputs([cls(color=torch.rand(2, 3))])


def test_scene_representation_glue_against_the_reference(ref):
    """The reference's HumanRF.density / forward and Decomposition4D.forward (humanrf.py:158-208, decomposition4d.py:124-135)
    executed live, with every tcnn module and the composition extension answering through the ORACLE's restatement of
    that one module.  What is compared is therefore the glue the oracle restates around them: frame -> segment routing,
    the +0.5 shift, local time, grid axis selection (xyz, xyt, yzt, xzt), composition call, truncated_exp * density_scale,
    geometry-feature slicing, (d+1)/2, camera embeddings while training / zeros otherwise.  The reference stores the
    composed features in fp16, hence the tolerances."""
    from helpers import positions_of, synthetic_rays
    from oracle import field as OF
    from oracle import hashgrid

    sizes, E = (6, 12, 6), 2
    frames = tuple(range(15, 15 + sum(sizes)))
    om = OF.make_model(sizes, frames, seed=5, table_init="trained", bf16=False, table_std=0.5, camera_embedding_dim=E)
    try:
        ref_model, _ = _import_reference_model()
        from humanrf_b200.synthetic import MODEL_KW

        theirs = ref_model.HumanRF(sorted_frame_numbers=frames, segment_sizes=sizes, **{**MODEL_KW, "camera_embedding_dim": E})
    finally:
        sys.modules.pop("tinycudann", None)
    import humanrf.scene_representation.decomposition4d as d4

    d4.Decomposition4D.to = lambda self, *a, **k: self                       # the reference parks idle segments on the CPU
    d4.tensor_composition_native.compose_tensors_forward = \
        lambda xyz, xyt, yzt, xzt, vectors, coords: OF.compose(xyz.float(), xyt.float(), yzt.float(), xzt.float(), vectors, coords).half()
    for s, fg in enumerate(theirs.feature_grids):
        with torch.no_grad():
            fg.vectors.copy_(om.segments[s].vectors)
        for k, name in enumerate(("xyz_encoding", "xyt_encoding", "yzt_encoding", "xzt_encoding")):
            enc = getattr(fg, name)
            enc.forward = (lambda x, s=s, k=k: hashgrid.encode(om.segments[s].grids[k], x.float(), om.segments[s].log2T).half())
    theirs.sigma_net.forward = lambda f: torch.relu(f.float() @ om.w_sigma[0].t()) @ om.w_sigma[1].t()

    def color_net(x):                                                        # [ (d+1)/2 | geo 15 | embedding E ] -> rgb
        d, rest = x[:, :3].float() * 2 - 1, x[:, 3:].float()
        inp = torch.cat((OF.sh4(d), rest, torch.ones((x.shape[0], 48 - 16 - rest.shape[1]))), dim=1)
        w1, w2, w3 = om.w_color
        h = torch.relu(torch.relu(inp @ w1.t()) @ w2.t())
        return torch.sigmoid((h @ w3.t())[:, :3])

    theirs.color_net.forward = color_net
    with torch.no_grad():
        theirs.camera_embeddings.weight.copy_(om.camera_embeddings)

    b = synthetic_rays(96, 24, frames, ragged=True, seed=12, n_distinct_frames=len(frames))
    pos, ri = positions_of(b), b["ri"]
    dirs, fr, cams = b["d"][ri], b["frames"][ri].view(-1, 1), b["cams"][ri].view(-1, 1)
    touched = set(om.f2s[b["frames"].numpy()].tolist())
    assert len(touched) == 3                                                 # all three segments are exercised
    for is_training in (True, False):
        q = ref.qio.QueryInput(is_training=is_training, positions=pos, directions=dirs, frame_numbers=fr,
                               unique_frame_numbers=torch.unique(b["frames"]).view(-1, 1), camera_numbers=cams)
        with torch.no_grad():
            out = theirs(q)
            dens = theirs.density(q)
            sigma, geo, rgb = om.forward(pos, dirs, fr.view(-1), cams.view(-1) if is_training else None)
        assert out.density.shape == sigma.shape and out.geometry_features.shape == (pos.shape[0], 15) and out.radiance.shape == rgb.shape
        assert torch.equal(out.density, dens.density)
        rel = ((out.density - sigma).abs() / sigma.abs().clamp_min(1e-3)).max().item()
        dg = (out.geometry_features.float() - geo).abs().max().item()
        dc = (out.radiance - rgb).abs().max().item()
        print(f"is_training={is_training}: density rel {rel:.2e}, geometry abs {dg:.2e}, radiance abs {dc:.2e}")
        assert rel < 2e-3 and dg < 2e-3 and dc < 1e-3       # measured 5e-5 / 5e-5 / 5e-6 (the fp16 feature buffer)
    # the embedding really is dropped at evaluation: the two passes differ in radiance only
    q_eval = ref.qio.QueryInput(is_training=False, positions=pos, directions=dirs, frame_numbers=fr,
                                unique_frame_numbers=torch.unique(b["frames"]).view(-1, 1), camera_numbers=cams)
    q_train = dataclasses.replace(q_eval, is_training=True)
    with torch.no_grad():
        assert (theirs(q_eval).radiance - theirs(q_train).radiance).abs().max() > 1e-3
