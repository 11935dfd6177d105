"""Oracle checks for SURVEY 8f-4 (visual-hull carving, adaptive temporal partitioning) on the CPU."""
from pathlib import Path

import numpy as np

from oracle import occupancy_tools as O
from scene import carve_scene, occupancy_sequence

GOLD = Path(__file__).parent / "golden" / "partitioning.npz"


def golden_cases():
    g = np.load(GOLD)
    for key in g.files:
        if key.endswith("_sizes"):
            name = key[:-6]
            n, speed, thr, seed = g[name + "_args"]
            yield name, int(n), (None if speed < 0 else float(speed)), float(thr), int(seed), g[key].tolist()


def test_partitioning_oracle_matches_reference_golden():
    """tests/golden/partitioning.npz holds the REFERENCE's own decisions (make_partitioning_golden.py)."""
    seen = 0
    for name, n, speed, thr, seed, sizes in golden_cases():
        grids = occupancy_sequence(n, speed, G=48, seed=seed)
        assert O.compute_adaptive_segment_sizes(grids, thr) == sizes, name
        seen += 1
    assert seen == 6


def test_segment_size_tables():
    assert [O.get_segment_size(k) for k in (6, 11, 12, 24, 25, 49, 50, 99, 100, 150)] == [6, 6, 12, 12, 25, 25, 50, 50, 100, 100]
    assert [O.get_final_segment_size(k) for k in (1, 6, 7, 12, 13, 26, 51, 100)] == [6, 6, 12, 12, 25, 50, 100, 100]


def test_carve_oracle_properties():
    sc = carve_scene(num_cameras=8, width=64, height=48)
    G, C = 24, 8
    full = O.generate_from_masks(sc["masks"], sc["projection_matrices"], sc["landscape"], 1, G, 64, 48)
    strict = O.generate_from_masks(sc["masks"], sc["projection_matrices"], sc["landscape"], C, G, 64, 48)
    mid = O.generate_from_masks(sc["masks"], sc["projection_matrices"], sc["landscape"], C // 2, G, 64, 48)
    assert set(np.unique(full)) <= {0, 255}
    # raising the coverage threshold can only carve more away
    assert ((strict == 255) <= (mid == 255)).all() and ((mid == 255) <= (full == 255)).all()
    assert 0 < (strict == 255).sum() < (full == 255).sum() < G ** 3
    # empty masks carve everything, threshold 0 is never "reached" by a miss and a hit reaches it at once
    none = O.generate_from_masks(np.zeros_like(sc["masks"]), sc["projection_matrices"], sc["landscape"], 1, G, 64, 48)
    assert (none == 0).all()
