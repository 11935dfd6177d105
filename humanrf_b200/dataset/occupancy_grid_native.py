"""occupancy_grid_native, mirroring the pybind class of actorshq/dataset/native/occupancy_grid.cu:8-95
(class name ``OccupanyGrid`` [sic], ``add_grid(grid) -> int``).  The returned int64 handle is the
device address of a bit-packed G^3 occupancy volume owned by the ring (the reference returns a
cudaTextureObject_t); ``ray_sampler_native`` interprets it consistently."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib as L


class OccupanyGrid:
    def __init__(self, grid_resolution: int, buffer_size: int):
        self._h = C.c_void_p()
        self.grid_resolution = int(grid_resolution)
        self.buffer_size = int(buffer_size)
        L.check(L.lib().hrf_occgrid_create(self.grid_resolution, self.buffer_size, C.byref(self._h)))

    def add_grid(self, grid: torch.Tensor) -> int:
        L.require_cuda(grid, "grid", torch.uint8)  # CHECK_CONTIGUITY_AND_DEVICE (occupancy_grid.cu:59)
        if grid.dim() != 3:
            raise RuntimeError("Provided grid doesn't have the correct resolution!")
        out = C.c_int64()
        L.check(L.lib().hrf_occgrid_add(self._h, grid.data_ptr(), grid.shape[0], grid.shape[1], grid.shape[2],
                                        L.stream(), C.byref(out)))
        return int(out.value)

    def lookup(self, handle: int, points: torch.Tensor) -> torch.Tensor:
        """Test hook: emulated ``tex3D(handle, p) > 0`` for normalised xyz points [N,3]."""
        p = L.require_cuda(points.contiguous(), "points", torch.float32)
        out = torch.empty(p.shape[0], dtype=torch.uint8, device=p.device)
        L.check(L.lib().hrf_occgrid_lookup(handle, self.grid_resolution, p.data_ptr(), p.shape[0], out.data_ptr(),
                                           L.stream()))
        return out.bool()

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                L.lib().hrf_occgrid_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass
