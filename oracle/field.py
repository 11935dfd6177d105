"""Oracle: HumanRF scene representation (4D decomposition + sigma / colour MLPs) on the CPU.

TEST INFRASTRUCTURE ONLY.  Follows
  humanrf/scene_representation/humanrf.py:79-109,158-208   (segment LUTs, density(), forward())
  humanrf/scene_representation/decomposition4d.py:73-135   (4 grids + vectors)
  humanrf/scene_representation/native/tensor_composition.cu:30-54  (vector lerp + composition)
  humanrf/utils/activation.py:6-39                         (truncated_exp)
and, PARITY UNPINNED (un-vendored tiny-cuda-nn): FullyFusedMLP layout/padding, the
Composite[SphericalHarmonics(4), Identity] encoding padded to 32 with 1.0.
The glue around the tcnn modules IS pinned: the reference's own HumanRF.__init__ / density /
forward and Decomposition4D.forward run on the CPU with each tcnn module answering through
this file's restatement of that module, and agree with OracleModel to 5e-5 (the reference's
fp16 feature buffer) -- tests/test_reference_live_cpu.py.

Everything is differentiable torch so autograd provides the backward oracle.
``bf16=True`` rounds at the same points as the CUDA kernels (tables, composed features,
hidden activations, geometry features, SH) while accumulating in fp32.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np
import torch

from . import hashgrid

PREDEFINED_SEGMENT_SIZES = (6, 12, 25, 50, 100)  # adaptive_temporal_partitioning.py:8


def rbf(x: torch.Tensor) -> torch.Tensor:
    """Round-trip through bf16 with a straight-through gradient."""
    return x + (x.to(torch.bfloat16).to(x.dtype) - x).detach()


class _TruncExp(torch.autograd.Function):
    """utils/activation.py:6-21 : exp forward, exp(clamp(x,-15,15)) backward."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return dy * torch.exp(x.clamp(-15, 15))


def truncated_exp(x):
    return _TruncExp.apply(x)


def segment_log2_hashmap_size(segment_size: int, log2_hashmap_size: int = 19) -> int:
    """humanrf.py:106-109."""
    return int(np.round(np.log2(segment_size / max(PREDEFINED_SEGMENT_SIZES) * (2 ** log2_hashmap_size))))


def frame_luts(sorted_frame_numbers, segment_sizes):
    """humanrf.py:79-103 : frame -> segment number, frame -> local_index / frames_in_segment."""
    num_frames = len(sorted_frame_numbers)
    end = np.cumsum(segment_sizes, dtype=np.int32)
    end[-1] = min(end[-1], num_frames)
    start = np.concatenate((np.zeros(1, np.int32), end[:-1]))
    f2s = np.full(sorted_frame_numbers[-1] + 1, -1, np.int32)
    f2t = np.full(sorted_frame_numbers[-1] + 1, -1, np.float32)
    for s in range(len(segment_sizes)):
        frames = [sorted_frame_numbers[j] for j in range(start[s], end[s])]
        for local, f in enumerate(frames):
            f2s[f] = s
            f2t[f] = local / len(frames)
    return f2s, f2t


def sh4(d: torch.Tensor) -> torch.Tensor:
    """Degree-4 real SH of x=2*in-1 where in=(dir+1)/2 (humanrf.py:192, tcnn spherical_harmonics.h)."""
    u = (d + 1) * 0.5
    v = u * 2 - 1
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    out = [
        torch.full_like(x, 0.28209479177387814),
        -0.48860251190291987 * y,
        0.48860251190291987 * z,
        -0.48860251190291987 * x,
        1.0925484305920792 * xy,
        -1.0925484305920792 * yz,
        0.94617469575755997 * z2 - 0.31539156525251999,
        -1.0925484305920792 * xz,
        0.54627421529603959 * x2 - 0.54627421529603959 * y2,
        0.59004358992664352 * y * (-3.0 * x2 + y2),
        2.8906114426405538 * xy * z,
        0.45704579946446572 * y * (1.0 - 5.0 * z2),
        0.3731763325901154 * z * (5.0 * z2 - 3.0),
        0.45704579946446572 * x * (1.0 - 5.0 * z2),
        1.4453057213202769 * z * (x2 - y2),
        0.59004358992664352 * x * (-x2 + 3.0 * y2),
    ]
    return torch.stack(out, dim=1)


def lerp_vectors(vectors: torch.Tensor, coords: torch.Tensor) -> torch.Tensor:
    """tensor_composition.cu:30-45.  vectors [4,VR,F]; coords [N,4] in [0,1] -> [4,N,F]."""
    vr = vectors.shape[1]
    out = []
    for i in range(4):
        c = coords[:, i] * vr - 0.5
        f = torch.floor(c)
        frac = (c - f).unsqueeze(1)
        i0 = torch.clamp(f, min=0.0).long()
        i1 = torch.clamp(f + 1, max=vr - 1).long()
        v0, v1 = vectors[i][i0], vectors[i][i1]
        out.append(v0 + frac * (v1 - v0))
    return torch.stack(out, 0)


def compose(xyz, xyt, yzt, xzt, vectors, coords):
    """tensor_composition.cu:49-52 : xyz*v_t + xyt*v_z + yzt*v_x + xzt*v_y."""
    sv = lerp_vectors(vectors, coords)
    return xyz * sv[3] + xyt * sv[2] + yzt * sv[0] + xzt * sv[1]


@dataclass
class OracleSegment:
    log2T: int
    grids: List[torch.Tensor]          # 4 x [entries, 2]  (xyz, xyt, yzt, xzt)
    vectors: torch.Tensor              # [4, 2048, 32]


@dataclass
class OracleModel:
    """Parameters in the reference's shapes.  MLP weights are row-major [out, in] (tcnn layout)."""
    segments: List[OracleSegment]
    f2s: np.ndarray
    f2t: np.ndarray
    w_sigma: List[torch.Tensor]        # [64,32], [16,64]
    w_color: List[torch.Tensor]        # [64,32], [64,64], [16,64]
    density_scale: float = 100.0
    bf16: bool = False
    camera_embeddings: torch.Tensor = None   # [160, E] or None (humanrf.py:75-76)
    sparse_grad: bool = False                # table gradients as sparse tensors (CPU training baseline only)

    def parameters(self):
        ps = []
        for s in self.segments:
            ps += list(s.grids) + [s.vectors]
        ps = ps + list(self.w_sigma) + list(self.w_color)
        return ps + ([self.camera_embeddings] if self.camera_embeddings is not None else [])

    def _q(self, x):
        return rbf(x) if self.bf16 else x

    def features(self, positions: torch.Tensor, frame_numbers: torch.Tensor) -> torch.Tensor:
        """humanrf.py:158-177 + decomposition4d.py:124-135.  positions in [-0.5,0.5]."""
        fr = frame_numbers.reshape(-1).long().numpy()
        seg = self.f2s[fr]
        tl = torch.from_numpy(self.f2t[fr]).unsqueeze(1)
        x = positions + 0.5
        xyzt = torch.cat((x, tl), dim=1)
        feats = torch.zeros((positions.shape[0], 32), dtype=positions.dtype)
        for s, sd in enumerate(self.segments):
            m = torch.from_numpy(np.nonzero(seg == s)[0])
            if m.numel() == 0:
                continue
            c = xyzt[m]
            g = [self._q(t) for t in sd.grids]
            sp = self.sparse_grad
            e = [hashgrid.encode(g[0], c[:, [0, 1, 2]], sd.log2T, sp),
                 hashgrid.encode(g[1], c[:, [0, 1, 3]], sd.log2T, sp),
                 hashgrid.encode(g[2], c[:, [1, 2, 3]], sd.log2T, sp),
                 hashgrid.encode(g[3], c[:, [0, 2, 3]], sd.log2T, sp)]
            feats = feats.index_add(0, m, compose(e[0], e[1], e[2], e[3], sd.vectors, c))
        return self._q(feats)

    def sigma_head(self, feats):
        """humanrf.py:181-186 : FullyFusedMLP 32->64 ReLU->16, density=trunc_exp(h0)*scale."""
        w1, w2 = [self._q(w) for w in self.w_sigma]
        h = self._q(torch.relu(feats @ w1.t()))
        o = h @ w2.t()
        return truncated_exp(o[:, 0]) * self.density_scale, o[:, 1:]

    def color_head(self, directions, geo, camera_numbers=None):
        """humanrf.py:188-206 : [SH16 | geo15 | camera embedding E (zeros at evaluation) | pad 1.0 to 32/48]
        -> 64 ReLU -> 64 ReLU -> 16 -> sigmoid[:3]."""
        n = directions.shape[0]
        parts = [self._q(sh4(directions)), self._q(geo)]
        if self.camera_embeddings is not None:
            E = self.camera_embeddings.shape[1]
            if camera_numbers is not None:
                parts.append(self._q(self.camera_embeddings[camera_numbers.reshape(-1).long()]))
            else:
                parts.append(torch.zeros((n, E), dtype=geo.dtype))
        width = self.w_color[0].shape[1]
        have = sum(p.shape[1] for p in parts)
        parts.append(torch.ones((n, width - have), dtype=geo.dtype))
        inp = torch.cat(parts, dim=1)
        w1, w2, w3 = [self._q(w) for w in self.w_color]
        h = self._q(torch.relu(inp @ w1.t()))
        h = self._q(torch.relu(h @ w2.t()))
        return torch.sigmoid((h @ w3.t())[:, :3])

    def density(self, positions, frame_numbers):
        return self.sigma_head(self.features(positions, frame_numbers))

    def forward(self, positions, directions, frame_numbers, camera_numbers=None):
        sigma, geo = self.density(positions, frame_numbers)
        return sigma, geo, self.color_head(directions, geo, camera_numbers)


def make_model(segment_sizes=(50,), sorted_frame_numbers=None, seed=123, table_init="trained",
               bf16=False, dtype=torch.float32, requires_grad=False, table_std=0.05,
               camera_embedding_dim=0) -> OracleModel:
    """Synthetic parameters per SURVEY 8(d): tables U(-1e-4,1e-4) ("tcnn") or N(0,table_std)
    ("trained"), vectors N(0,0.1) (decomposition4d.py:76-78), Xavier-uniform MLP weights."""
    g = torch.Generator().manual_seed(seed)
    if sorted_frame_numbers is None:
        sorted_frame_numbers = tuple(range(15, 15 + sum(segment_sizes)))
    f2s, f2t = frame_luts(sorted_frame_numbers, segment_sizes)
    segs = []
    for ss in segment_sizes:
        l2 = segment_log2_hashmap_size(ss)
        total = hashgrid.level_table(l2)[5]
        grids = []
        for _ in range(4):
            if table_init == "tcnn":
                t = (torch.rand((total, 2), generator=g) * 2 - 1) * 1e-4
            else:
                t = torch.randn((total, 2), generator=g) * table_std
            grids.append(t.to(dtype))
        vec = (torch.randn((4, 2048, 32), generator=g) * 0.1).to(dtype)
        segs.append(OracleSegment(l2, grids, vec))

    def xavier(o, i):
        a = float(np.sqrt(6.0 / (i + o)))
        return ((torch.rand((o, i), generator=g) * 2 - 1) * a).to(dtype)

    k = 32 if camera_embedding_dim == 0 else 48
    m = OracleModel(segs, f2s, f2t, [xavier(64, 32), xavier(16, 64)],
                    [xavier(64, k), xavier(64, 64), xavier(16, 64)], 100.0, bf16)
    if camera_embedding_dim > 0:
        m.camera_embeddings = torch.randn((160, camera_embedding_dim), generator=g).to(dtype)   # nn.Embedding init N(0,1)
    if requires_grad:
        for p in m.parameters():
            p.requires_grad_(True)
    return m
