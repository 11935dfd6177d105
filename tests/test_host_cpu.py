"""CPU suite: host logic and oracle against the golden vectors generated from the reference
(tests/golden/make_golden.py), layout tables, and the C-ABI export list."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from humanrf_b200.dataset.cameras import inverse_kr, projection_matrix_world2pixel
from humanrf_b200.dataset.input_batch import InputBatch
from humanrf_b200.input import merge_input_batches
from humanrf_b200.scene_representation.grid_layout import GridLayout, mlp_blob_permutation, segment_log2_hashmap_size
from humanrf_b200.utils.activation import truncated_exp
from humanrf_b200.utils.loss import bce_loss
from oracle import field as ofield
from oracle import hashgrid, rendering

ROOT = Path(__file__).resolve().parent.parent
G = np.load(ROOT / "tests/golden/reference_host.npz")
FIELDS = ["ray_origins", "ray_directions", "minmaxes", "rgba", "ray_masks", "frame_numbers", "unique_frame_numbers",
          "camera_numbers", "sample_distances", "ray_indices"]


@pytest.mark.parametrize("case", ["a", "b", "c", "d"])
def test_merge_input_batches_matches_reference(case):
    nb = int(G[f"merge_{case}_nb"])
    budget = int(G[f"merge_{case}_budget"])
    batches = []
    for bi in range(nb):
        kw = {f: torch.from_numpy(G[f"merge_{case}_in{bi}_{f}"]) for f in FIELDS}
        batches.append(InputBatch(width=64, height=48, **kw))
    out = merge_input_batches(batches, None if budget < 0 else budget)
    for f in FIELDS:
        got = getattr(out, f)
        if f == "unique_frame_numbers":
            got = torch.sort(got.reshape(-1))[0]
        exp = G[f"merge_{case}_out_{f}"]
        assert got.dtype == torch.from_numpy(exp).dtype, f
        np.testing.assert_array_equal(got.numpy(), exp, err_msg=f)
    assert out.width == 64 and out.height == 48


def test_truncated_exp_matches_reference():
    for impl in (truncated_exp, ofield.truncated_exp):
        x = torch.from_numpy(G["texp_x"]).clone().requires_grad_(True)
        y = impl(x)
        y.backward(torch.from_numpy(G["texp_dy"]))
        np.testing.assert_array_equal(y.detach().numpy(), G["texp_y"])
        np.testing.assert_array_equal(x.grad.numpy(), G["texp_dx"])


def test_bce_loss_matches_reference():
    for impl in (bce_loss, rendering.bce_loss):
        out = impl(torch.from_numpy(G["bce_pred"]), torch.from_numpy(G["bce_target"]))
        np.testing.assert_array_equal(out.numpy(), G["bce_out"])


def test_inverse_kr_matches_reference_camera():
    for i, cp in enumerate(G["cam_params"]):
        w2p = projection_matrix_world2pixel(1028, 752, cp[:3], cp[3:], np.array([1.773863, 1.773863 * 1028 / 752]),
                                            np.array([0.5, 0.5]))
        np.testing.assert_allclose(np.linalg.inv(w2p), G["cam_world2pixel_inv_full"][i], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(inverse_kr(w2p), G["cam_inverse_krs"][i], rtol=2e-6, atol=1e-7)


def test_grid_layout_matches_oracle_and_survey():
    # entries per grid for log2T = 19..15
    expected = {19: 6984576, 18: 3695768, 17: 1947288, 16: 1015808, 15: 524288}
    for l2, total in expected.items():
        lay = GridLayout(l2)
        sc, rs, of, sz, hs, tot = hashgrid.level_table(l2)
        assert lay.n_entries == tot == total
        np.testing.assert_array_equal(lay.scale, sc)
        np.testing.assert_array_equal(lay.res, rs)
        np.testing.assert_array_equal(lay.offset, of)
        np.testing.assert_array_equal(lay.size, sz)
        assert lay.hashed_mask == sum(1 << i for i, h in enumerate(hs) if h)
    lay = GridLayout(19)
    assert list(lay.res[:4]) == [32, 43, 56, 74] and lay.hashed_mask == 0xFFF0
    np.testing.assert_allclose(lay.scale[[5, 10, 15]], [127.000015, 511.00018, 2047.0015], rtol=1e-6)
    assert [segment_log2_hashmap_size(s, 19) for s in (6, 12, 25, 50, 100)] == [15, 16, 17, 18, 19]
    assert [ofield.segment_log2_hashmap_size(s) for s in (6, 12, 25, 50, 100)] == [15, 16, 17, 18, 19]


def test_mlp_blob_permutation():
    for emb, n in ((0, 10240), (2, 11264)):
        dst, src = mlp_blob_permutation(emb)
        assert sorted(src.tolist()) == list(range(n)) and len(set(dst.tolist())) == n and dst.max() < 22528 // 2
    dst, src = mlp_blob_permutation(0)
    # element (n=9, k=17) of the first layer [64,32]: core (kg=2, ng=1), row 1, col 1
    assert src[dst.tolist().index((2 * 8 + 1) * 64 + 1 * 8 + 1)] == 9 * 32 + 17
    # colour W2 follows colour W1 [64,K]: byte 6144 + 128*K
    dst48, src48 = mlp_blob_permutation(2)
    assert dst48[src48.tolist().index(3072 + 64 * 48)] == (6144 + 128 * 48) // 2
    assert dst[src.tolist().index(3072 + 64 * 32)] == (6144 + 128 * 32) // 2


def test_c_abi_exports_every_declared_symbol():
    from humanrf_b200 import _lib

    header = (ROOT / "include/humanrf_b200.h").read_text()
    declared = sorted(set(re.findall(r"\b(hrf_[a-z0-9_]+)\s*\(", header)))
    assert declared == _lib.exported_symbols()
    if not _lib.LIB_PATH.exists():
        pytest.skip("library not built yet (python -m humanrf_b200.build)")
    handle = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        assert hasattr(handle, name), name
    assert _lib.lib().hrf_version() >= 1


def test_product_does_not_import_oracle():
    for p in (ROOT / "humanrf_b200").rglob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p


def test_dropin_maps_reference_import_paths():
    import sys

    ref = "/root/reference"
    if not Path(ref).exists():
        pytest.skip("reference checkout not present (GPU box)")
    import humanrf_b200.dropin as dropin

    sys.path.insert(0, ref)
    try:
        dropin.install()
        from humanrf.scene_representation.humanrf import HumanRF as A
        from humanrf_b200.scene_representation.humanrf import HumanRF as B
        import humanrf.volume_rendering as vr
        import actorshq.dataset.ray_sampler_native as rs

        assert A is B and hasattr(vr, "prune_samples") and hasattr(rs, "get_samples_occupancy_minmax")
    finally:
        sys.path.remove(ref)
        for k in list(sys.modules):
            if k.startswith(("humanrf.", "actorshq.")) or k in ("humanrf", "actorshq"):
                del sys.modules[k]
