"""Native training step for the HumanRF hot path (the fast path behind bench.py --mode train and the
data-parallel driver).  Same maths as the reference's Trainer.train_step (humanrf/trainer.py:229-255:
random background, Huber(delta=0.01) + 1e-3 * BCE, Adam lr 1e-2 betas (0.9,0.99) eps 1e-15,
lr * lr_decay^(min(step/max,1)), run.py:101-104), but:

* gradients are written by the fused backward kernels straight into ONE flat fp32 bucket (no per-parameter
  zero-filled tensors, no autograd graph), laid out grid-major so that under data parallelism the bucket
  is reduced in 5 NCCL messages, each overlapping the scatter of the next table and the Adam of the previous one;
* Adam is one fused kernel per parameter that also refreshes the bf16 shadow table the forward reads;
* bf16 needs no GradScaler, so the inf-check host sync of trainer.py:250-252 disappears.

The autograd-compatible route (humanrf_b200.volume_rendering.render + torch.optim.Adam) stays available for
running the reference's trainer.py unchanged.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L
from .parallel import (active_segments, allreduce_bucket_, allreduce_spans, grid_major_bucket_layout,
                       union_batch_loss_scale)
from .scene_representation.humanrf import HumanRF
from .volume_rendering import ray_offsets


class FusedTrainer:
    def __init__(self, model: HumanRF, lr: float = 1e-2, betas=(0.9, 0.99), eps: float = 1e-15, lr_decay: float = 0.5,
                 max_steps: int = 50001, bce_loss_weight: float = 1e-3, huber_delta: float = 0.01,
                 render_step_size: float = 4e-4, world_size: int = 1, process_group=None, prune: bool = True,
                 seed: int = 123, overlap_allreduce: bool = True):
        self.model, self.lr, self.betas, self.eps = model, lr, betas, eps
        self.lr_decay, self.max_steps = lr_decay, max_steps
        self.bce_w, self.delta, self.step_size = bce_loss_weight, huber_delta, render_step_size
        self.world, self.pg, self.prune = world_size, process_group, prune
        self.overlap_allreduce = overlap_allreduce
        self.params: List[torch.nn.Parameter] = model.hot_parameters()
        dev = self.params[0].device
        m = model
        S = m.num_segments
        # Bucket layout, GRID-MAJOR: [grid 0 of every segment | grid 1 ... | grid 2 ... | grid 3 ... | vectors of every
        # segment, MLPs, camera embeddings].  Region k is complete as soon as the scatter launch of grid k has run, so under
        # data parallelism its all-reduce overlaps the scatter of grid k+1 (hot_parameters() itself is segment-major).
        self.slices, self.regions, self.adam_order = grid_major_bucket_layout([p.numel() for p in self.params], S)
        total = self.regions[-1][1]
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)       # the all-reduce bucket
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad_views = [self.grad[a:b] for a, b in self.slices]
        self.t = 0                                  # optimiser steps taken (drives the learning-rate schedule)
        self.steps = [0] * len(self.params)         # torch.optim.Adam's per-parameter state['step'] (bias corrections)
        self.gen = torch.Generator(device=dev).manual_seed(seed)
        self.nat = model.native()
        sg = (L.SegmentGrads * S)()
        for s_ in range(S):
            for k in range(4):
                sg[s_].grid[k] = self.grad_views[5 * s_ + k].data_ptr()
            sg[s_].vectors = self.grad_views[5 * s_ + 4].data_ptr()
        self.sg_dev = torch.from_numpy(np.frombuffer(bytes(sg), dtype=np.uint8).copy()).to(dev)
        i = 5 * S
        self.mlp_grad = self.grad[self.slices[i][0]:self.slices[i + 1][1]]   # sigma params then colour params, contiguous
        assert self.mlp_grad.numel() == model.mlp_grad_elems
        self.emb_grad = self.grad_views[i + 2] if model.camera_embedding_dim > 0 else None
        self.last = {}
        self.profile = False

    # ----------------------------------------------------------------------------------------------
    def current_lr(self) -> float:
        """LambdaLR of run.py:102-104 evaluated at the number of COMPLETED steps: the scheduler is stepped after the
        optimiser (trainer.py:251-253), so the first update runs at the full learning rate."""
        return self.lr * self.lr_decay ** min(self.t / self.max_steps, 1.0)

    def step(self, o, d, frames, t, ri, rgba, num_rays: int, kernel_event=None, return_loss: bool = False,
             background: Optional[torch.Tensor] = None, cameras: Optional[torch.Tensor] = None, bwd_events=None):
        """One optimisation step on a ray batch given in InputBatch layout (device tensors).
        Returns the number of kernels launched, or the loss value when return_loss."""
        lib, nat, dev = L.lib(), self.model.native(), t.device
        launches = 0
        marks = [] if self.profile else None

        def mark(name):
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((name, e))

        mark("start")
        step = self.step_size
        t = t.reshape(-1)
        # Segments this batch touches (humanrf.py:162-179): the reference gives the others no gradient, so Adam leaves
        # their parameters, moments and step counters alone.  Decided on the device; read back with the prune counter.
        S = self.model.num_segments
        used = None
        if S > 1:
            used = active_segments(self.model.frame_numbers_to_segment_numbers, frames, S).to(torch.int64)
            if self.world > 1:                       # a segment is active if any rank's batch touches it
                dist.all_reduce(used, op=dist.ReduceOp.MAX, group=self.pg)
        active = None
        # ---- prune_samples (volume_rendering.py:42-84): jitter, density-only pass, visibility compaction
        if self.prune:
            t = t + torch.rand(t.shape, device=dev, generator=self.gen) * step
            n0 = t.shape[0]
            off0 = ray_offsets(ri, num_rays)
            sigma0 = nat.density_early_stop(nat.samples_rays(o, d, frames, t, ri), off0, num_rays, step)
            keep = torch.empty(n0, dtype=torch.uint8, device=dev)
            kept_off = torch.empty(num_rays + 1, dtype=torch.int32, device=dev)
            t2 = torch.empty(n0, dtype=torch.float32, device=dev)
            ri2 = torch.empty(n0, dtype=torch.int64, device=dev)
            counter = torch.zeros(1, dtype=torch.int64, device=dev)
            L.check(lib.hrf_prune(sigma0.data_ptr(), t.data_ptr(), ri.data_ptr(), off0.data_ptr(), num_rays, step, 1e-4,
                                  1e-4, keep.data_ptr(), kept_off.data_ptr(), t2.data_ptr(), ri2.data_ptr(),
                                  counter.data_ptr(), L.stream()))
            mark("prune_enqueued")
            if used is None:
                kept = int(counter.item())
            else:                                    # one read for both
                host = torch.cat((counter, used)).cpu().tolist()
                kept, active = int(host[0]), [bool(x) for x in host[1:]]
            t, ri = t2[:kept], ri2[:kept]
            launches += 9
            mark("prune_synced")
            off = kept_off                                   # hrf_prune's scan IS the ray-offset table of the survivors
        if used is not None and active is None:
            active = [bool(x) for x in used.cpu().tolist()]
        n = t.shape[0]
        # ---- forward: fused field + compositing
        samples = nat.samples_rays(o, d, frames, t, ri, cameras if self.model.camera_embedding_dim > 0 else None)
        sigma, _, rgb, feat = nat.forward(samples, 1, want_geo=False, want_feat=True)
        if kernel_event is not None:
            kernel_event.record()
        mark("forward")
        if not self.prune:
            off = ray_offsets(ri, num_rays)
        bg = background if background is not None else torch.rand((num_rays, 3), device=dev, generator=self.gen)  # trainer.py:237
        color = torch.empty((num_rays, 3), dtype=torch.float32, device=dev)
        wsum = torch.empty((num_rays, 1), dtype=torch.float32, device=dev)
        L.check(lib.hrf_composite_forward(sigma.data_ptr(), rgb.data_ptr(), t.data_ptr(), off.data_ptr(), num_rays, step,
                                          bg.data_ptr(), color.data_ptr(), wsum.data_ptr(), None, L.stream()))
        # ---- loss on the [R,3] outputs (tiny): Huber + BCE, gradients by autograd on the leaf outputs
        color.requires_grad_(True)
        wsum.requires_grad_(True)
        mask = rgba[:, 3:4]
        gt = rgba[:, :3] * mask + bg * (1 - mask)
        photo = torch.nn.functional.huber_loss(color, gt, delta=self.delta, reduction="mean")
        pc = torch.clamp(wsum, min=0, max=1)
        bce = -(mask * torch.log(pc + 1e-10) + (1 - mask) * torch.log(1 - pc + 1e-10))
        loss = photo + bce.mean() * self.bce_w
        # data parallel: weight by this rank's share of the union batch (humanrf_b200/parallel.py)
        (loss * union_batch_loss_scale(num_rays, dev, self.pg) if self.world > 1 else loss).backward()
        mark("composite+loss")
        # ---- backward: compositing, then the fused field backward into the flat bucket
        d_sigma = torch.empty(n, dtype=torch.float32, device=dev)
        d_rgb = torch.empty((n, 3), dtype=torch.float32, device=dev)
        L.check(lib.hrf_composite_backward(sigma.data_ptr(), rgb.data_ptr(), t.data_ptr(), off.data_ptr(), num_rays, step,
                                           bg.data_ptr(), color.grad.data_ptr(), wsum.grad.reshape(-1).data_ptr(),
                                           d_sigma.data_ptr(), d_rgb.data_ptr(), L.stream()))
        self.grad.zero_()
        if bwd_events is not None:
            bwd_events[0].record()
        ws = torch.empty(n * 40, dtype=torch.float32, device=dev)   # 160 B / sample
        L.check(lib.hrf_field_backward_mlp(C.byref(nat.field), C.byref(samples), d_sigma.data_ptr(), d_rgb.data_ptr(),
                                           feat.data_ptr(), self.mlp_grad.data_ptr(), L.ptr(self.emb_grad), ws.data_ptr(),
                                           L.stream()))
        egrid = feat.data_ptr() + 64 * n
        works = None
        # data parallel: only the gradients of the segments this step touched (on any rank) are reduced
        spans = allreduce_spans(self.slices, self.regions, S, active) if self.world > 1 else None
        if self.world == 1 or not self.overlap_allreduce:
            L.check(lib.hrf_field_backward_tables(C.byref(nat.field), C.byref(samples), self.sg_dev.data_ptr(), egrid,
                                                  ws.data_ptr(), 0, 4, L.stream()))
            if self.world > 1 and active is None:
                allreduce_bucket_(self.grad, self.pg)                       # everything is active: one message
            elif self.world > 1:
                for region in spans:
                    for a, b in region:
                        allreduce_bucket_(self.grad[a:b], self.pg)
            launches += 8 + 12 + ((1 if active is None else sum(len(r) for r in spans)) if self.world > 1 else 0)
        else:
            # table k's gradient region is reduced (NCCL, its own stream) while table k+1 is still being scattered; the
            # sum's mean over ranks is folded into Adam's grad_scale
            works = []
            for k in range(4):
                L.check(lib.hrf_field_backward_tables(C.byref(nat.field), C.byref(samples), self.sg_dev.data_ptr(), egrid,
                                                      ws.data_ptr(), k, 1, L.stream()))
                works.append([allreduce_bucket_(self.grad[a:b], self.pg, async_op=True) for a, b in spans[k]])
            works.append([allreduce_bucket_(self.grad[a:b], self.pg, async_op=True) for a, b in spans[4]])
            launches += 8 + 15 + sum(len(r) for r in spans)
        if bwd_events is not None:
            bwd_events[1].record()
        mark("backward")
        self.apply_adam(1.0 / self.world, works, active)
        launches += len(self.params) + 3
        mark("allreduce+adam")
        self.last = {"samples": n, "loss": loss.detach()}
        if marks is not None:
            torch.cuda.synchronize()
            self.last["phases_ms"] = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(marks[:-1], marks[1:])}
        if return_loss:
            return float(loss.item())
        return launches

    def apply_adam(self, grad_scale: float, works=None, active=None) -> None:
        """Adam over the bucket in region order; `works` = the pending all-reduces of each of the 5 regions (a list per
        region), waited for just before the first parameter of that region.  `active` (bool per segment, None = all) selects the segments
        that took part in this step: the others are skipped entirely and keep their own step counters, exactly what
        torch.optim.Adam does with parameters whose .grad is None (the reference's trainer.py:174,251)."""
        nat = self.nat
        lr = self.current_lr()
        self.t += 1
        S = self.model.num_segments
        with torch.no_grad():
            for j, i in enumerate(self.adam_order):
                region = min(j // S, 4)
                if works is not None and works[region]:
                    for w in works[region]:
                        w.wait()
                    works[region] = []
                if active is not None and i < 5 * S and not active[i // 5]:
                    continue
                self.steps[i] += 1
                shadow = nat.shadows[i // 5][i % 5] if (i < 5 * S and i % 5 < 4) else None
                self._adam(i, shadow, lr, grad_scale)
            nat.repack_mlp()

    def _adam(self, i: int, shadow: Optional[torch.Tensor], lr: float, grad_scale: float) -> None:
        p = self.params[i]
        a, b = self.slices[i]
        L.check(L.lib().hrf_adam_step(p.data_ptr(), self.exp_avg[a:b].data_ptr(), self.exp_avg_sq[a:b].data_ptr(),
                                      self.grad[a:b].data_ptr(), L.ptr(shadow), b - a, lr, self.betas[0], self.betas[1],
                                      self.eps, self.steps[i], grad_scale, L.stream()))
