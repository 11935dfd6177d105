"""Pins the integer emulation of the texture unit's 1.8 fixed-point trilinear '> 0' test (csrc/sampler.cu
occ_lookup, oracle/sampler.py tex_occupied) against a REAL tex3D configured as the reference does
(occupancy_grid.cu:17-38), through oracle/tex_probe.cu."""
import ctypes
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import sampler as S
from scene import ellipsoid_grid

pytestmark = pytest.mark.gpu
PROBE = Path(__file__).resolve().parent.parent / "oracle/_build/libtexprobe.so"


def _hw(cuda, grid, pts):
    if not PROBE.exists():
        pytest.skip("oracle/_build/libtexprobe.so not built (python -c 'import __graft_entry__ as g; g.build()')")
    lib = ctypes.CDLL(str(PROBE))
    lib.tex_probe.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
    g = torch.from_numpy(grid).to(cuda).contiguous()
    p = torch.from_numpy(pts).to(cuda).contiguous()
    out = torch.empty(p.shape[0], device=cuda)
    torch.cuda.synchronize()
    assert lib.tex_probe(g.data_ptr(), grid.shape[0], p.data_ptr(), p.shape[0], out.data_ptr()) == 0
    return out.cpu().numpy()


@pytest.mark.parametrize("G", [64, 256])
def test_emulation_matches_hardware(cuda, G):
    from humanrf_b200.dataset.occupancy_grid_native import OccupanyGrid

    rng = np.random.default_rng(G)
    grid = ellipsoid_grid(G, 1)
    single = np.zeros((G, G, G), np.uint8)
    single[G // 2, G // 3, G // 4] = 255
    for gi, gr in enumerate((grid, single)):
        pts = rng.uniform(-0.02, 1.02, (400000, 3)).astype(np.float32)
        # adversarial: points a few float ulps around voxel-centre planes of occupied voxels and 1/256 sub-steps
        zz, yy, xx = np.nonzero(gr)
        sel = rng.integers(0, zz.size, 200000)
        base = (np.stack([xx[sel], yy[sel], zz[sel]], 1) + 0.5) / G
        sub = rng.integers(-300, 301, base.shape) / (256.0 * G)
        jit = rng.integers(-3, 4, base.shape) * np.spacing(np.float32(base))
        pts2 = (base + sub + jit).astype(np.float32)
        allp = np.concatenate([pts, pts2])
        hw = _hw(cuda, gr, allp) > 0
        emu = S.tex_occupied(gr, allp)
        og = OccupanyGrid(G, 1)
        h = og.add_grid(torch.from_numpy(gr).to(cuda))
        dev = og.lookup(h, torch.from_numpy(allp).to(cuda)).cpu().numpy()
        np.testing.assert_array_equal(dev, emu, err_msg="CUDA emulation != oracle emulation")
        mism = np.nonzero(hw != emu)[0]
        print(f"G={G} grid#{gi}: {mism.size} hardware/emulation mismatches out of {allp.shape[0]}")
        if mism.size:
            for i in mism[:10]:
                print("  p*G=", allp[i] * G, "hw", hw[i], "emu", emu[i])
        assert mism.size <= 3e-4 * allp.shape[0]   # exact except at .5 ties of the hardware's internal roundings
