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
    """Data parallelism over scenes: flat fp32 gradient buckets, NCCL all-reduce overlapped with the backward pass.

    Parameters are laid out in the flat buffer in REVERSE registration order (the order in which the backward pass produces
    their gradients: output conv and decoder first, patch embed last) and cut into `n_buckets` contiguous buckets of similar
    size.  `arm()` before the last backward of a step installs nothing new - the post-accumulate hooks registered once at
    construction count arrivals; when the last gradient of a bucket has been accumulated, that bucket is packed with one
    multi-tensor copy kernel and `all_reduce(async_op=True)` is issued, so the collective of the decoder's gradients runs on
    NCCL's stream while the encoder's backward is still computing.  `finish()` waits for the outstanding works and returns the
    flat buffer (sum over ranks); the 1/world_size is folded into the optimizer kernel through `grad_scale`.  `reduce()` is the
    non-overlapped form (pack everything, one collective per bucket) for callers that run backward themselves."""

    def __init__(self, params: Iterable[torch.nn.Parameter], process_group=None, n_buckets: int = 4):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        order = list(range(len(self.params)))[::-1]            # backward order
        numels = [self.params[i].numel() for i in order]
        total = int(sum(numels))
        starts = np.concatenate([[0], np.cumsum(numels)[:-1]]).astype(np.int64) if numels else np.zeros(0, np.int64)
        # offsets[i] = position of parameter i (registration order) in the flat buffer
        self.offsets = np.zeros(len(self.params), dtype=np.int64)
        for pos, i in enumerate(order):
            self.offsets[i] = starts[pos]
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.world = torch.distributed.get_world_size(process_group) if torch.distributed.is_initialized() else 1
        # buckets: contiguous ranges of the backward-ordered parameter list
        n_buckets = max(1, min(n_buckets, len(order)))
        target = total / n_buckets
        self.buckets = []                                      # (param indices, flat start, flat end)
        cur, cur_start, acc = [], 0, 0
        for pos, i in enumerate(order):
            cur.append(i)
            acc += numels[pos]
            if acc >= target * (len(self.buckets) + 1) and len(self.buckets) < n_buckets - 1:
                self.buckets.append((cur, cur_start, acc))
                cur, cur_start = [], acc
        if cur:
            self.buckets.append((cur, cur_start, acc))
        self._bucket_of = {}
        for b, (idx, _, _) in enumerate(self.buckets):
            for i in idx:
                self._bucket_of[i] = b
        self.plans = [_ChunkPlan([self.params[i].numel() for i in idx]) if dev.type == "cuda" else None for idx, _, _ in self.buckets]
        self._armed = False
        self._pending = [0] * len(self.buckets)
        self._works = []
        for i, p in enumerate(self.params):
            p.register_post_accumulate_grad_hook(self._make_hook(i))

    # ------------------------------------------------------------------ bucket operations
    def views(self):
        return [self.flat[o:o + p.numel()].view_as(p) for o, p in zip(self.offsets, self.params)]

    def _pack_bucket(self, b: int):
        idx, start, end = self.buckets[b]
        dev = self.flat.device
        grads = [self.params[i].grad for i in idx]
        if any(g is None for g in grads):
            raise RuntimeError("GradAllReducer: a parameter has no gradient (unused parameter?)")
        if dev.type == "cuda":
            d_ptr = self.flat.data_ptr() + self.offsets[np.asarray(idx)] * 4
            tbl = self.plans[b].table(None, _ptrs(grads), None, None, d_ptr, dev)
            call("nmae_multi_copy", tbl, self.plans[b].n, device=dev)
        else:  # gloo / CPU tests of the host logic
            for i, g in zip(idx, grads):
                o = int(self.offsets[i])
                self.flat[o:o + g.numel()].copy_(g.reshape(-1))

    def _launch_bucket(self, b: int):
        self._pack_bucket(b)
        if self.world > 1:
            _, start, end = self.buckets[b]
            self._works.append(torch.distributed.all_reduce(self.flat[start:end], group=self.group, async_op=True))

    def _make_hook(self, i: int):
        def hook(_param):
            if not self._armed:
                return
            b = self._bucket_of[i]
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._launch_bucket(b)
        return hook

    # ------------------------------------------------------------------ public
    def arm(self):
        """Call before the LAST backward of an optimiser step: buckets are reduced as soon as their gradients are complete."""
        self._armed = True
        self._pending = [len(idx) for idx, _, _ in self.buckets]
        self._works = []

    def finish(self) -> torch.Tensor:
        """After the backward armed by arm(): wait for the collectives (stream-level) and return the flat buffer."""
        if not self._armed:
            raise RuntimeError("GradAllReducer.finish() without arm()")
        self._armed = False
        for b, left in enumerate(self._pending):
            if left != 0:                                   # a hook did not fire (e.g. gradient produced outside autograd)
                self._launch_bucket(b)
        for w in self._works:
            w.wait()
        self._works = []
        return self.flat

    def reduce(self) -> torch.Tensor:
        """Non-overlapped form: pack every param.grad, all-reduce bucket by bucket, return the flat buffer (sum over ranks)."""
        self._armed = False
        self._works = []
        for b in range(len(self.buckets)):
            self._launch_bucket(b)
        for w in self._works:
            w.wait()
        self._works = []
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

    # ------------------------------------------------------------------ (de)serialisation
    def state_dict(self):
        """torch.optim.AdamW-compatible state: every parameter's entry carries 'step' (the bias-correction count) next to
        'exp_avg' / 'exp_avg_sq', so the dict loads into torch.optim.AdamW and vice versa."""
        for gi, group in enumerate(self.param_groups):
            if gi in self._plans:
                for p in self._plans[gi][0]:
                    self.state[p]["step"] = torch.tensor(float(self._steps))
        return super().state_dict()

    def load_state_dict(self, state_dict):
        """Restores the moments AND the step count (from the per-parameter 'step' entries torch.optim.AdamW also writes); the flat
        moment buffers are rebuilt from the loaded tensors on the next step()."""
        super().load_state_dict(state_dict)
        self._plans = {}
        steps = [int(float(st["step"])) for st in self.state.values() if "step" in st]
        if steps:
            self._steps = max(steps)

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
        # the parameters changed behind torch's version counters: rebuild every cached tensor-core weight blob (one launch)
        from .functional import weight_blobs
        weight_blobs.refresh()
        return None
