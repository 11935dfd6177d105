"""merge_input_batches, mirroring humanrf/input.py:10-55 (same semantics, including the
sample-budget cut-off at ray granularity and the recomputed unique frame numbers)."""
from __future__ import annotations

from typing import List, Optional

import torch

from .dataset.input_batch import InputBatch


def merge_input_batches(input_batches: List[InputBatch], max_num_samples: Optional[int] = None) -> InputBatch:
    merged = InputBatch()
    first = input_batches[0]
    for key, val in vars(first).items():
        if key == "ray_indices":
            continue
        if val is None:
            setval = None
        elif isinstance(val, torch.Tensor):
            setval = torch.cat([getattr(b, key) for b in input_batches], dim=0)
        elif isinstance(val, int):
            setval = val
        else:
            raise RuntimeError("Unknown data type in the input_batches!")
        setattr(merged, key, setval)

    if first.ray_indices is not None:
        # rebase ray indices by the number of rays of the preceding batches (input.py:24-31)
        parts, base = [], 0
        for b in input_batches:
            parts.append(b.ray_indices + base)
            base += b.num_rays
        merged.ray_indices = torch.cat(parts, dim=0)

    if max_num_samples is not None:
        num_rays, num_samples = merged.num_rays, merged.num_samples
        if num_samples > max_num_samples:
            cutoff = merged.ray_indices[max_num_samples]  # first ray that no longer fits (input.py:37)
            sample_keep = merged.ray_indices < cutoff
            for key, val in list(vars(merged).items()):
                if not isinstance(val, torch.Tensor):
                    continue
                if key == "ray_masks":
                    setval = val[val.cumsum(0) < cutoff]
                elif val.shape[0] == num_rays:
                    setval = val[:cutoff]
                elif val.shape[0] == num_samples:
                    setval = val[sample_keep]
                else:
                    continue  # e.g. unique_frame_numbers: recomputed below
                setattr(merged, key, setval)

    if merged.frame_numbers is not None:
        merged.unique_frame_numbers = torch.unique(merged.frame_numbers, sorted=False, return_inverse=False).view(-1, 1)
    return merged
