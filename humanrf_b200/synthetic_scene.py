"""Synthetic ActorsHQ-like scene for the sampler tests and the bench: occupancy grid =
union of 3 ellipsoids in [-0.5,0.5]^3, cameras on a ring at distance 2 looking at the origin."""
from __future__ import annotations

import numpy as np

from humanrf_b200.dataset.cameras import inverse_kr, projection_matrix_world2pixel


def ellipsoid_grid(G=128, seed=0):
    rng = np.random.default_rng(seed)
    c = (np.arange(G) + 0.5) / G - 0.5
    z, y, x = np.meshgrid(c, c, c, indexing="ij")
    occ = np.zeros((G, G, G), bool)
    for _ in range(3):
        ctr = rng.uniform(-0.15, 0.15, 3)
        rad = rng.uniform(0.08, 0.25, 3)
        occ |= ((x - ctr[0]) / rad[0]) ** 2 + ((y - ctr[1]) / rad[1]) ** 2 + ((z - ctr[2]) / rad[2]) ** 2 <= 1.0
    return (occ * 255).astype(np.uint8)


def look_at_camera(pos, width, height, f=1.773863):
    """RDF camera at `pos` looking at the origin; returns (world2pixel 4x4, rotation axis-angle)."""
    fwd = -pos / np.linalg.norm(pos)
    up = np.array([0.0, -1.0, 0.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    Rm = np.stack([right, down, fwd], 1)  # camera->world
    # axis-angle from rotation matrix
    ang = np.arccos(np.clip((np.trace(Rm) - 1) / 2, -1, 1))
    if ang < 1e-9:
        aa = np.zeros(3)
    else:
        aa = ang / (2 * np.sin(ang)) * np.array([Rm[2, 1] - Rm[1, 2], Rm[0, 2] - Rm[2, 0], Rm[1, 0] - Rm[0, 1]])
    return projection_matrix_world2pixel(width, height, aa, pos, np.array([f, f * width / height]), np.array([0.5, 0.5]))


def make_scene(num_images=4, width=96, height=72, G=128, seed=0, portrait_every=3):
    rng = np.random.default_rng(seed)
    inv, org, land = [], [], []
    for i in range(num_images):
        a = 2 * np.pi * i / num_images + 0.3
        pos = np.array([2.0 * np.cos(a), rng.uniform(-0.4, 0.4), 2.0 * np.sin(a)])
        ls = not (portrait_every and i % portrait_every == portrait_every - 1)
        w, h = (width, height) if ls else (height, width)
        inv.append(inverse_kr(look_at_camera(pos, w, h)))
        org.append(pos.astype(np.float32))
        land.append(ls)
    grids = [ellipsoid_grid(G, seed + i % 2) for i in range(num_images)]
    P = num_images * width * height
    return dict(
        inverse_krs=np.stack(inv).astype(np.float32), camera_origins=np.stack(org).astype(np.float32),
        landscape=np.array(land), grids=grids, G=G, width=width, height=height,
        frame_numbers=(15 + np.arange(num_images) % 6).astype(np.int32),
        camera_numbers=rng.integers(0, 160, num_images).astype(np.int32),
        rgba=rng.integers(0, 256, (P, 4)).astype(np.uint8), light_mask=rng.random(P) < 0.05,
        aabb=np.array([[-0.45, -0.5, -0.4], [0.5, 0.45, 0.5]], np.float32))


class _Cam:
    def __init__(self, width, height, rot, translation, f=1.773863):
        self.width, self.height = width, height
        self.rotation_axisangle, self.translation = np.asarray(rot, float), np.asarray(translation, float)
        self.focal_length, self.principal_point = np.array([f, f * width / height]), np.array([0.5, 0.5])

    def projection_matrix_world2pixel(self):
        return projection_matrix_world2pixel(self.width, self.height, self.rotation_axisangle, self.translation,
                                             self.focal_length, self.principal_point)


class SyntheticDataset:
    """Duck-typed stand-in for actorshq.dataset.volumetric_dataset.VolumetricDataset (the methods DataLoader calls)."""

    def __init__(self, num_cameras=5, frames=range(15, 21), width=64, height=48, G=64, seed=0):
        import copy

        self._copy = copy
        rng = np.random.default_rng(seed)
        self.frames = list(frames)
        self.cameras = []
        for i in range(num_cameras):
            a = 2 * np.pi * i / num_cameras + 0.2
            pos = np.array([3.0 * np.cos(a), rng.uniform(-0.5, 0.5), 3.0 * np.sin(a)]) + np.array([0.3, 1.0, -0.2])
            w2p = look_at_camera(pos - np.array([0.3, 1.0, -0.2]), width, height)   # reuse the rotation of a centred camera
            fwd = -(pos - np.array([0.3, 1.0, -0.2])); fwd /= np.linalg.norm(fwd)
            up = np.array([0.0, -1.0, 0.0]); right = np.cross(fwd, up); right /= np.linalg.norm(right); down = np.cross(fwd, right)
            Rm = np.stack([right, down, fwd], 1)
            ang = np.arccos(np.clip((np.trace(Rm) - 1) / 2, -1, 1))
            aa = ang / (2 * np.sin(ang)) * np.array([Rm[2, 1] - Rm[1, 2], Rm[0, 2] - Rm[2, 0], Rm[1, 0] - Rm[0, 1]])
            self.cameras.append(_Cam(width, height, aa, pos))
        self.aabb = np.array([[-0.45, 0.25, -0.95], [1.05, 1.75, 0.55]])       # world-space box (extent 1.5) around (0.3,1,-0.2)
        self._grids = {f: ellipsoid_grid(G, seed + k % 2) for k, f in enumerate(self.frames)}
        self._rng_seed = seed

    def get_aabb(self, frame_numbers=None):
        return self.aabb

    def get_scaled_cameras(self, scene_offset, scene_scale):
        cams = self._copy.deepcopy(self.cameras)
        for c in cams:
            c.translation = (c.translation + scene_offset) * scene_scale
        return cams

    def get_rgb(self, camera_number, frame_number, normalize=True):
        rng = np.random.default_rng(self._rng_seed * 7919 + camera_number * 131 + frame_number)
        c = self.cameras[camera_number]
        return rng.integers(0, 256, (c.height, c.width, 3)).astype(np.float32) / np.float32(255)

    def get_mask(self, camera_number, frame_number, normalize=True):
        rng = np.random.default_rng(self._rng_seed * 104729 + camera_number * 17 + frame_number)
        c = self.cameras[camera_number]
        return (rng.random((c.height, c.width, 1)) > 0.4).astype(np.float32)

    def get_occupancy_grid(self, frame_number):
        return self._grids[frame_number]


def occupancy_sequence(n, speed, G=48, seed=0):
    """n occupancy grids of a drifting ellipsoid pair (speed = drift per frame; None = static with sudden jumps), the
    input of adaptive temporal partitioning."""
    rng = np.random.default_rng(seed)
    c = (np.arange(G) + 0.5) / G - 0.5
    z, y, x = np.meshgrid(c, c, c, indexing="ij")
    ctr = rng.uniform(-0.1, 0.1, (2, 3))
    rad = rng.uniform(0.1, 0.2, (2, 3))
    vel = rng.normal(size=(2, 3))
    vel /= np.linalg.norm(vel, axis=1, keepdims=True)
    grids = []
    for f in range(n):
        occ = np.zeros((G, G, G), bool)
        for k in range(2):
            if speed is None:
                p = ctr[k] + 0.12 * vel[k] * ((f // 37) % 3)
            else:
                p = ctr[k] + speed * f * vel[k] * np.cos(0.05 * f)
            occ |= ((x - p[0]) / rad[k, 0]) ** 2 + ((y - p[1]) / rad[k, 1]) ** 2 + ((z - p[2]) / rad[k, 2]) ** 2 <= 1.0
        grids.append((occ * 255).astype(np.uint8))
    return grids


def carve_scene(num_cameras=12, width=96, height=72, seed=0, portrait_every=4):
    """Foreground masks of an ellipsoid union seen from a camera ring + the transposed world2pixel matrices, the input
    of generate_from_masks (generate_occupancy_grids_from_masks.py:44-93)."""
    rng = np.random.default_rng(seed)
    ctr = rng.uniform(-0.12, 0.12, (3, 3))
    rad = rng.uniform(0.08, 0.22, (3, 3))
    masks = np.zeros((num_cameras, width * height), np.uint8)
    mats, land = [], []
    for i in range(num_cameras):
        a = 2 * np.pi * i / num_cameras + 0.2
        pos = np.array([2.0 * np.cos(a), rng.uniform(-0.5, 0.5), 2.0 * np.sin(a)])
        ls = not (portrait_every and i % portrait_every == portrait_every - 1)
        w, h = (width, height) if ls else (height, width)
        M = look_at_camera(pos, w, h)
        ikr = inverse_kr(M)
        u, v = np.meshgrid(np.arange(w) + 0.5, np.arange(h) + 0.5)
        d = np.stack([u, v, np.ones_like(u)], -1) @ np.asarray(ikr, np.float64)      # inverse_kr is stored transposed
        d /= np.linalg.norm(d, axis=-1, keepdims=True)
        hit = np.zeros((h, w), bool)
        for k in range(3):                                      # ray / ellipsoid intersection
            o_, d_ = (pos - ctr[k]) / rad[k], d / rad[k]
            A, B, Cc = (d_ * d_).sum(-1), 2 * (d_ * o_).sum(-1), (o_ * o_).sum() - 1
            hit |= B * B - 4 * A * Cc >= 0
        masks[i] = (hit * 255).astype(np.uint8).reshape(-1)
        mats.append(np.asarray(M, np.float32).T.copy())          # transposed, as the reference passes them to GLM
        land.append(ls)
    return dict(masks=masks, projection_matrices=np.stack(mats), landscape=np.array(land), width=width, height=height)
