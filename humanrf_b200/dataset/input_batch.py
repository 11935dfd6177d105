"""InputBatch, mirroring actorshq/dataset/input_batch.py:8-50 field for field."""
from __future__ import annotations

from dataclasses import dataclass

import torch


@dataclass
class InputBatch:
    # (#rays x 3): [torch.float]
    ray_origins: torch.Tensor = None
    # (#rays x 3): [torch.float]
    ray_directions: torch.Tensor = None
    # (#rays x 2): [torch.float]
    minmaxes: torch.Tensor = None
    # (#rays x 4): [torch.float]
    rgba: torch.Tensor = None
    # (>=#rays x 1): [torch.bool]  -- original batch size, False where the sampler dropped the ray
    ray_masks: torch.Tensor = None
    # (#rays x 1): [torch.int32]
    frame_numbers: torch.Tensor = None
    # (K x 1): [torch.int32]
    unique_frame_numbers: torch.Tensor = None
    # (#rays x 1): [torch.int32]
    camera_numbers: torch.Tensor = None
    # (#samples x 1): [torch.float]
    sample_distances: torch.Tensor = None
    # (#samples): [torch.int64], sorted ascending (ray-major)
    ray_indices: torch.Tensor = None
    width: int = None
    height: int = None

    @property
    def num_rays(self) -> int:
        return self.ray_origins.shape[0]

    @property
    def num_samples(self) -> int:
        return self.sample_distances.shape[0]
