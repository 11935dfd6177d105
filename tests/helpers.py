"""Shared builders for the parity tests: identical parameters in the oracle model and the
CUDA-backed HumanRF module, and the seeded synthetic ray batch of SURVEY 8(d)."""
from __future__ import annotations

import numpy as np
import torch

from oracle import field as ofield

from humanrf_b200.synthetic import MODEL_KW, input_batch_of, synthetic_rays  # noqa: F401,E402


def make_pair(segment_sizes=(6,), seed=123, table_std=6.0, device="cuda", bf16=True, first_frame=15, cam_emb=0):
    """Returns (oracle_model, humanrf_module) holding the same parameter values."""
    from humanrf_b200.scene_representation.humanrf import HumanRF

    frames = tuple(range(first_frame, first_frame + sum(segment_sizes)))
    om = ofield.make_model(segment_sizes, frames, seed=seed, table_init="trained", bf16=bf16, table_std=table_std,
                           camera_embedding_dim=cam_emb)
    m = HumanRF(sorted_frame_numbers=frames, segment_sizes=tuple(segment_sizes), **{**MODEL_KW, "camera_embedding_dim": cam_emb})
    with torch.no_grad():
        for s, fg in enumerate(m.feature_grids):
            for k, g in enumerate(fg.grids()):
                g.copy_(om.segments[s].grids[k].reshape(-1))
            fg.vectors.copy_(om.segments[s].vectors)
        m.sigma_net.params.copy_(torch.cat([w.reshape(-1) for w in om.w_sigma]))
        m.color_net.params.copy_(torch.cat([w.reshape(-1) for w in om.w_color]))
        if cam_emb:
            m.camera_embeddings.weight.copy_(om.camera_embeddings)
    return om, m.to(device), frames


def positions_of(b):
    """volume_rendering.py:66-69 in fp32 (mul then add)."""
    return b["o"][b["ri"]] + b["t"].unsqueeze(1) * b["d"][b["ri"]]


def rel_err(a, b, floor=1e-6):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)
