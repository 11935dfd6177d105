"""tensor_composition_native, mirroring the pybind functions of
humanrf/scene_representation/native/tensor_composition.cu:120-225 (half features, fp32 vectors).
The fused field kernels do this composition internally; this stand-alone op exists so code written
against the reference extension (decomposition4d.py:8-39) keeps working and so the first-party CUDA
of the reference can be diffed against ours on the GPU box."""
from __future__ import annotations

from typing import List

import torch

from .. import _lib as L


def _check(ts):
    for i, t in enumerate(ts):
        L.require_cuda(t, f"arg{i}")


def compose_tensors_forward(xyz_features, xyt_features, yzt_features, xzt_features, xyzt_vectors,
                            xyzt_coordinates) -> torch.Tensor:
    _check((xyz_features, xyt_features, yzt_features, xzt_features, xyzt_vectors, xyzt_coordinates))
    n, f = xyz_features.shape
    out = torch.empty_like(xyz_features)
    L.check(L.lib().hrf_compose_tensors_forward(
        xyz_features.data_ptr(), xyt_features.data_ptr(), yzt_features.data_ptr(), xzt_features.data_ptr(),
        xyzt_vectors.data_ptr(), xyzt_coordinates.data_ptr(), n, f, xyzt_vectors.shape[1], out.data_ptr(), L.stream()))
    return out


def compose_tensors_backward(xyz_features, xyt_features, yzt_features, xzt_features, xyzt_vectors, xyzt_coordinates,
                             d_output_features) -> List[torch.Tensor]:
    _check((xyz_features, xyt_features, yzt_features, xzt_features, xyzt_vectors, xyzt_coordinates, d_output_features))
    n, f = xyz_features.shape
    d = [torch.empty_like(xyz_features) for _ in range(4)]
    dv = torch.empty_like(xyzt_vectors)
    L.check(L.lib().hrf_compose_tensors_backward(
        xyz_features.data_ptr(), xyt_features.data_ptr(), yzt_features.data_ptr(), xzt_features.data_ptr(),
        xyzt_vectors.data_ptr(), xyzt_coordinates.data_ptr(), d_output_features.data_ptr(), n, f,
        xyzt_vectors.shape[1], d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), dv.data_ptr(),
        L.stream()))
    return d + [dv]
