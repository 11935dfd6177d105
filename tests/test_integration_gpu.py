"""End-to-end: DataLoader -> prune -> merge -> FusedTrainer -> tile renderer -> PSNR on a held-out camera of a
teacher-rendered synthetic dataset (examples/train_synthetic.py).  PSNR must rise."""
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu


def test_training_raises_held_out_psnr(cuda):
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "examples"))
    import train_synthetic

    hist = train_synthetic.main(steps=150, log_every=50, quiet=True)
    print("PSNR history", hist)
    assert hist[-1][1] > hist[0][1] + 3.0, hist
