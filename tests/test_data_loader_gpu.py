"""DataLoader with the GPU-resident pool (humanrf_b200/dataset/data_loader.py) against the reference's __next__
contract (data_loader.py:531-660) and the oracle sampler on the very rays it drew."""
import numpy as np
import pytest
import torch

from oracle import sampler as S
from scene import SyntheticDataset

pytestmark = pytest.mark.gpu


def _loader(mode, **kw):
    from humanrf_b200.dataset.data_loader import DataLoader as DL

    ds = SyntheticDataset()
    args = dict(dataset=ds, device="cuda", mode=mode, dataloader_output_mode=DL.OutputMode.RAYS_AND_SAMPLES,
                space_pruning_mode=DL.SpacePruningMode.OCCUPANCY_GRID, batch_size=1024, camera_numbers=(0, 1, 2, 3, 4),
                frame_numbers=tuple(ds.frames), max_buffer_size=12)
    args.update(kw)
    return DL(**args), ds, DL


def test_training_batches_match_oracle_on_the_drawn_rays(cuda):
    np.random.seed(0); torch.manual_seed(0)
    from humanrf_b200.dataset.data_loader import DataLoader as DLC

    dl, ds, DL = _loader(DLC.Mode.TRAINING, max_num_frames_per_batch=3, use_mask=True, filter_light_bloom=False)
    assert dl.buffer_size == 10 and dl.occupancy_grids_buffer_size == 3      # 5 cameras x (3-1) frames (data_loader.py:251-256)
    it = iter(dl)
    for step in range(4):
        dl.batch_size = 1024 - 100 * step                                    # the trainer mutates batch_size (trainer.py:156-161)
        b = next(it)
        assert len(set(dl.frame_numbers_cuda.tolist())) <= 3                 # max_num_frames_per_batch bounds the pool
        idx = dl.last_ray_indices.cpu().numpy()
        assert idx.shape[0] == dl.batch_size and b.ray_masks.shape == (dl.batch_size, 1)
        slots = dl.buffer_size
        grids = []
        frames = dl.frame_numbers_cuda.cpu().numpy()
        for s in range(slots):
            grids.append(ds.get_occupancy_grid(int(frames[s])))
        exp = S.get_data(dl.pixel_colors.view(-1, 4).cpu().numpy(), dl.light_mask.view(-1).cpu().numpy(), frames,
                         dl.camera_numbers_cuda.cpu().numpy(), grids, dl.landscape_mode_cuda.cpu().numpy(), idx,
                         dl.inverse_krs_cuda.cpu().numpy(), dl.camera_origins_cuda.cpu().numpy(), dl.aabb.cpu().numpy(),
                         dl.occupancy_grid_resolution, 64, 48, 4e-4, False)
        got = [b.ray_origins, b.ray_directions, b.rgba, b.frame_numbers.view(-1), b.camera_numbers.view(-1), b.minmaxes,
               b.ray_masks.view(-1), b.sample_distances.view(-1), b.ray_indices]
        for g, e in zip(got, exp):
            np.testing.assert_array_equal(g.cpu().numpy(), e)
        assert b.ray_indices.dtype == torch.int64 and b.frame_numbers.dtype == torch.int32 and b.rgba.dtype == torch.float32
        assert sorted(b.unique_frame_numbers.view(-1).tolist()) == sorted(set(b.frame_numbers.view(-1).tolist()))
        assert b.width == 64 and b.height == 48
    # scene normalisation puts the AABB inside [-0.5, 0.5]^3 (data_loader.py:182-215)
    assert dl.aabb.min() >= -0.5 - 1e-6 and dl.aabb.max() <= 0.5 + 1e-6 and abs(float((dl.aabb[1] - dl.aabb[0]).max()) - 1.0) < 1e-5


def test_validation_covers_every_pixel_once_and_test_has_no_rgba(cuda):
    from humanrf_b200.dataset.data_loader import DataLoader as DL

    seq = [(0, 15), (3, 17), (1, 15)]
    for mode in (DL.Mode.VALIDATION, DL.Mode.TEST):
        kw = dict(render_sequence=seq, batch_size=1000, max_buffer_size=2)
        if mode == DL.Mode.VALIDATION:
            kw.update(use_mask=True, filter_light_bloom=False)
        dl, ds, _ = _loader(mode, **kw)
        assert dl.num_camera_frame_pairs == 3 and dl.num_batches_per_full_image == 4 and len(dl) == 3 * 64 * 48
        masks, frames, n = [], [], 0
        for b in dl:
            masks.append(b.ray_masks.view(-1).cpu())
            frames += b.frame_numbers.view(-1).tolist()
            assert (b.rgba is None) == (mode == DL.Mode.TEST)
            n += 1
        assert n == 3 * 4 and torch.cat(masks).numel() == 3 * 64 * 48
        per_image = torch.cat(masks).view(3, -1).sum(1)
        assert (per_image > 0).all()
        assert set(frames) == {15, 17}
    with pytest.raises(RuntimeError, match="render_sequence"):
        _loader(DL.Mode.TRAINING, max_num_frames_per_batch=2, use_mask=True, filter_light_bloom=False, render_sequence=seq)


def test_validation_second_pass_reloads_the_ring(cuda):
    """The trainer re-iterates the validation loader every N steps with max_buffer_size=1 (run.py): after a pass over a
    render sequence longer than the pool, the slots hold the LAST images; the second pass must render the same images
    as the first (colours, cameras, occupancy handles), not stale slots."""
    from humanrf_b200.dataset.data_loader import DataLoader as DL

    seq = [(0, 15), (3, 17), (1, 16)]
    for max_buffer in (1, 2):
        dl, ds, _ = _loader(DL.Mode.VALIDATION, render_sequence=seq, batch_size=64 * 48, max_buffer_size=max_buffer,
                            use_mask=True, filter_light_bloom=False)
        passes = []
        for _ in range(2):
            passes.append([(b.rgba.clone(), b.ray_origins.clone(), b.ray_directions.clone(), b.sample_distances.clone(),
                            b.frame_numbers.clone(), b.camera_numbers.clone()) for b in dl])
        assert len(passes[0]) == len(passes[1]) == 3
        for first, second in zip(*passes):
            for a, b in zip(first, second):
                assert torch.equal(a, b)
        # and the three images differ from each other (so a stale slot would have been noticed)
        assert not torch.equal(passes[0][0][1][0], passes[0][1][1][0])
