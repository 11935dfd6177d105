"""occupancy_grid_generation_native.generate_from_masks, mirroring the pybind function of
actorshq/toolbox/native/occupancy_grid_generation.cu:83-125 (visual-hull carving from foreground masks)."""
import torch

from .. import _lib as L


def generate_from_masks(masks: torch.Tensor, projection_matrices: torch.Tensor, landscape_modes: torch.Tensor,
                        camera_coverage_threshold: int, grid_resolution: int, width: int, height: int) -> torch.Tensor:
    if masks.size(1) != width * height:
        raise RuntimeError("The number mask entries per camera has to be equal to width*height!")
    L.require_cuda(masks, "masks", torch.uint8)
    L.require_cuda(projection_matrices, "projection_matrices", torch.float32)
    L.require_cuda(landscape_modes, "landscape_modes", torch.bool)
    grid = torch.empty((grid_resolution,) * 3, dtype=torch.uint8, device=masks.device)
    L.check(L.lib().hrf_occupancy_from_masks(masks.data_ptr(), projection_matrices.data_ptr(), landscape_modes.data_ptr(),
                                             projection_matrices.size(0), int(camera_coverage_threshold),
                                             int(grid_resolution), int(width), int(height), grid.data_ptr(), L.stream()))
    return grid
