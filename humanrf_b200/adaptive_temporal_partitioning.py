"""Adaptive temporal partitioning on the GPU, mirroring humanrf/adaptive_temporal_partitioning.py:8-107 (same function
name, arguments and decisions).  The union of occupancy grids (eq. 2) is a bit-packed volume on the device and the
occupied-voxel counts (eq. 3) come from one fused OR + popcount kernel per frame (hrf_occupancy_union_count); the
reference ORs and counts 16 MB uint8 numpy arrays on the CPU."""
from typing import List

import numpy as np
import torch

from . import _lib as L

PREDEFINED_SEGMENT_SIZES = [6, 12, 25, 50, 100]


def get_segment_size(num_frames: int):
    for idx, segment_size in enumerate(PREDEFINED_SEGMENT_SIZES[:-1]):
        if num_frames < PREDEFINED_SEGMENT_SIZES[idx + 1]:
            return segment_size
    return PREDEFINED_SEGMENT_SIZES[-1]


def get_final_segment_size(num_frames_left: int):
    for segment_size in PREDEFINED_SEGMENT_SIZES:
        if num_frames_left <= segment_size:
            return segment_size


class _Cluster:
    def __init__(self, num_voxels: int, device):
        self.bits = torch.zeros((num_voxels + 31) // 32, dtype=torch.int32, device=device)
        self.count = torch.zeros(1, dtype=torch.int64, device=device)
        self.num_frames = 0

    def add_grid(self, grid_u8: torch.Tensor) -> int:
        """eq. (2) + eq. (3): OR the grid into the union and return the union's occupied-voxel count."""
        L.check(L.lib().hrf_occupancy_union_count(self.bits.data_ptr(), grid_u8.data_ptr(), grid_u8.numel(),
                                                  self.count.data_ptr(), L.stream()))
        self.num_frames += 1
        return int(self.count.item())


def compute_adaptive_segment_sizes(dataset, sorted_frame_numbers: List[int], expansion_factor_threshold: float = 1.25,
                                   device="cuda") -> List[int]:
    min_segment_size, max_segment_size = min(PREDEFINED_SEGMENT_SIZES), max(PREDEFINED_SEGMENT_SIZES)
    cluster = None
    segment_sizes = []
    fnum_idx, total, decided = 0, len(sorted_frame_numbers), 0
    initial_occupancy = 0
    while fnum_idx < total:
        grid = torch.from_numpy(dataset.get_occupancy_grid(frame_number=sorted_frame_numbers[fnum_idx])).to(device).contiguous()
        if cluster is None:
            cluster = _Cluster(grid.numel(), device)
            union_occupancy = initial_occupancy = cluster.add_grid(grid)      # the first grid alone
        else:
            union_occupancy = cluster.add_grid(grid)
        if cluster.num_frames >= min_segment_size:
            with np.errstate(all="ignore"):                                  # an empty first grid divides by zero, as numpy does
                expansion_factor = np.float64(union_occupancy) / np.float64(initial_occupancy)   # eq. (4)
            if expansion_factor > expansion_factor_threshold or cluster.num_frames >= max_segment_size:
                segment_size = get_segment_size(cluster.num_frames)
                decided += segment_size
                cluster = None
                fnum_idx = decided
                segment_sizes.append(segment_size)
                continue
        fnum_idx += 1
    if decided < total:
        segment_sizes.append(get_final_segment_size(total - decided))
    assert sum(segment_sizes) >= total
    return segment_sizes
