"""Camera math the sampler needs, mirroring actorshq/dataset/camera_data.py:50-102 and the
inverse-KR table of actorshq/dataset/data_loader.py:182-215 (scene normalisation + transposed
inverse projection, float32)."""
from __future__ import annotations

import numpy as np


def rotation_from_axisangle(r: np.ndarray) -> np.ndarray:
    """Rodrigues formula (scipy Rotation.from_rotvec(...).as_matrix() equivalent)."""
    r = np.asarray(r, np.float64)
    theta = np.linalg.norm(r)
    if theta < 1e-12:
        return np.eye(3)
    k = r / theta
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(theta) * K + (1 - np.cos(theta)) * (K @ K)


def projection_matrix_world2pixel(width, height, rotation_axisangle, translation, focal_length, principal_point):
    """camera_data.py:64-102."""
    intr = np.array([[width * focal_length[0], 0, width * principal_point[0]],
                     [0, height * focal_length[1], height * principal_point[1]], [0, 0, 1.0]])
    c2w = np.eye(4)
    c2w[:3, :3] = rotation_from_axisangle(rotation_axisangle)
    c2w[:3, 3] = translation
    w2p = np.eye(4)
    w2p[:3] = intr @ np.linalg.inv(c2w)[:3]
    return w2p


def inverse_kr(world2pixel: np.ndarray) -> np.ndarray:
    """data_loader.py:194-207 : inv(world2pixel)[:3,:3], transposed (GLM column-major), float32."""
    return np.linalg.inv(world2pixel)[:3, :3].T.astype(np.float32)


def normalise_scene(aabb: np.ndarray):
    """data_loader.py:182-184 : offset = -mean(aabb), scale = 1/max(extent)."""
    aabb = np.asarray(aabb, np.float64)
    return -aabb.mean(0), 1.0 / np.max(aabb[1] - aabb[0])
