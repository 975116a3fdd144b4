"""Fused gradient-clip + AdamW step and the data-parallel gradient all-reduce
(reference: run_swin_mae3d.py:355-357 DDP, :588-598 AdamW/OneCycleLR, :663-669 clip + step).

`FusedAdamWClip` is a torch.optim.Optimizer (so OneCycleLR can cycle `lr` and `betas` exactly as it does with
torch.optim.AdamW) whose step is three kernel launches in total: sum of squared gradients, then
clip-coefficient + AdamW update, over a multi-tensor chunk table - no host synchronisation.
"""
from __future__ import annotations

from typing import Iterable, Optional

import numpy as np
import torch

from ._lib import call

CHUNK = 1 << 16


class _ChunkPlan:
    """Static split of a parameter list into <=CHUNK-element pieces; only base pointers change per step."""

    def __init__(self, numels):
        pidx, off, cnt = [], [], []
        for i, n in enumerate(numels):
            for o in range(0, n, CHUNK):
                pidx.append(i)
                off.append(o)
                cnt.append(min(CHUNK, n - o))
        self.pidx = np.asarray(pidx, dtype=np.int64)
        self.off_bytes = np.asarray(off, dtype=np.int64) * 4
        self.cnt = np.asarray(cnt, dtype=np.int64)
        self.n = len(pidx)
        self._np = np.zeros((max(self.n, 1), 6), dtype=np.int64)
        self._last = None
        self._dev_table = None

    def table(self, p_ptr, g_ptr, m_ptr, v_ptr, d_ptr, device):
        t = self._np
        for col, base in enumerate((p_ptr, g_ptr, m_ptr, v_ptr)):
            t[:self.n, col] = (base[self.pidx] + self.off_bytes) if base is not None else 0
        t[:self.n, 4] = self.cnt
        t[:self.n, 5] = (d_ptr[self.pidx] + self.off_bytes) if d_ptr is not None else 0
        # gradient buffers usually come back at the same addresses every step (caching allocator): re-upload
        # only when something moved.  A fresh pinned staging tensor per upload keeps the async copy race-free.
        if self._last is None or not np.array_equal(t, self._last):
            self._last = t.copy()
            self._dev_table = torch.from_numpy(self._last).pin_memory().to(device, non_blocking=True)
        return self._dev_table


def _ptrs(tensors):
    return np.fromiter((t.data_ptr() for t in tensors), dtype=np.int64, count=len(tensors))


class GradAllReducer:
    """Data parallelism over scenes: one flat fp32 bucket, one NCCL all-reduce per step.

    pack (one multi-tensor copy kernel) -> dist.all_reduce(SUM) on the flat bucket; the 1/world_size is folded
    into the optimizer kernel through `grad_scale`."""

    def __init__(self, params: Iterable[torch.nn.Parameter], process_group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        numels = [p.numel() for p in self.params]
        self.offsets = np.concatenate([[0], np.cumsum(numels)[:-1]]).astype(np.int64) if numels else np.zeros(0, np.int64)
        dev = self.params[0].device
        self.flat = torch.zeros(int(sum(numels)), dtype=torch.float32, device=dev)
        self.plan = _ChunkPlan(numels) if dev.type == "cuda" else None
        self.world = torch.distributed.get_world_size(process_group) if torch.distributed.is_initialized() else 1

    def views(self):
        return [self.flat[o:o + p.numel()].view_as(p) for o, p in zip(self.offsets, self.params)]

    def reduce(self) -> torch.Tensor:
        """Pack every param.grad into the flat bucket, all-reduce it, return the bucket (sum over ranks)."""
        dev = self.flat.device
        if dev.type == "cuda":
            grads = [p.grad for p in self.params]
            if any(g is None for g in grads):
                raise RuntimeError("GradAllReducer.reduce(): a parameter has no gradient (unused parameter?)")
            d_ptr = self.flat.data_ptr() + self.offsets * 4
            tbl = self.plan.table(None, _ptrs(grads), None, None, d_ptr, dev)
            call("nmae_multi_copy", tbl, self.plan.n, device=dev)
        else:  # gloo / CPU tests of the host logic
            for v, p in zip(self.views(), self.params):
                v.copy_(p.grad)
        if self.world > 1:
            torch.distributed.all_reduce(self.flat, group=self.group)
        return self.flat


class FusedAdamWClip(torch.optim.Optimizer):
    """torch.nn.utils.clip_grad_norm_(params, clip) + torch.optim.AdamW.step() in two kernels."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, clip_grad_norm: float = 0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.clip = float(clip_grad_norm)
        self._plans = {}
        self._steps = 0
        self.norm_sq: Optional[torch.Tensor] = None

    def _group_state(self, gi, group):
        if gi not in self._plans:
            ps = [p for p in group["params"] if p.requires_grad]
            numels = [p.numel() for p in ps]
            dev = ps[0].device
            tot = int(sum(numels))
            m = torch.zeros(tot, dtype=torch.float32, device=dev)
            v = torch.zeros(tot, dtype=torch.float32, device=dev)
            offs = np.concatenate([[0], np.cumsum(numels)[:-1]]).astype(np.int64)
            for p, o in zip(ps, offs):     # expose the moments under torch.optim.AdamW's state names
                mv, vv = m[o:o + p.numel()].view_as(p), v[o:o + p.numel()].view_as(p)
                st = self.state[p]
                if "exp_avg" in st:        # moments restored by load_state_dict(): move them into the flat buffers
                    mv.copy_(st["exp_avg"])
                    vv.copy_(st["exp_avg_sq"])
                st["exp_avg"], st["exp_avg_sq"] = mv, vv
            self._plans[gi] = (ps, _ChunkPlan(numels), m, v, offs)
        return self._plans[gi]

    def grad_norm(self) -> torch.Tensor:
        """Global L2 norm of the last step's (scaled) gradients, as a device scalar (no sync)."""
        return self.norm_sq.sqrt().float() * self._last_scale

    @torch.no_grad()
    def step(self, closure=None, flat_grads: Optional[torch.Tensor] = None, flat_offsets=None, grad_scale: float = 1.0):
        if closure is not None:
            raise ValueError("FusedAdamWClip does not support closures")
        self._steps += 1
        t = self._steps
        self._last_scale = grad_scale
        tables = []
        dev = None
        for gi, group in enumerate(self.param_groups):
            ps, plan, m, v, offs = self._group_state(gi, group)
            dev = ps[0].device
            if flat_grads is not None:
                if len(self.param_groups) != 1:
                    raise ValueError("flat gradient buckets need a single param group")
                g_ptr = flat_grads.data_ptr() + np.asarray(flat_offsets, dtype=np.int64) * 4
            else:
                grads = [p.grad for p in ps]
                if any(g is None for g in grads):
                    raise RuntimeError("FusedAdamWClip.step(): a parameter has no gradient")
                if any(not g.is_contiguous() for g in grads):
                    raise RuntimeError("FusedAdamWClip.step(): non-contiguous gradient")
                g_ptr = _ptrs(grads)
            tbl = plan.table(_ptrs(ps), g_ptr, m.data_ptr() + offs * 4, v.data_ptr() + offs * 4, None, dev)
            tables.append((group, plan, tbl))
        if self.norm_sq is None:
            self.norm_sq = torch.zeros(1, dtype=torch.float64, device=dev)
        if len(tables) == 1:
            call("nmae_multi_sumsq", tables[0][2], tables[0][1].n, self.norm_sq, device=dev)
        else:  # several groups: one global norm
            acc = torch.zeros_like(self.norm_sq)
            for _, plan, tbl in tables:
                call("nmae_multi_sumsq", tbl, plan.n, self.norm_sq, device=dev)
                acc += self.norm_sq
            self.norm_sq.copy_(acc)
        for group, plan, tbl in tables:
            b1, b2 = group["betas"]
            call("nmae_adamw_clip_step", tbl, plan.n, self.norm_sq, self.clip, float(grad_scale), float(group["lr"]),
                 float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]), 1.0 - b1 ** t, 1.0 - b2 ** t,
                 device=dev)
        return None
