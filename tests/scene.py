"""Synthetic scenes live in the package (humanrf_b200/synthetic_scene.py: the bench and the examples use them too)."""
from humanrf_b200.synthetic_scene import *  # noqa: F401,F403
from humanrf_b200.synthetic_scene import _Cam  # noqa: F401
