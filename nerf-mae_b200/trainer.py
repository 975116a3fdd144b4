"""One MAE pretraining step as the reference driver runs it (run_swin_mae3d.py:644-709 `Trainer.train_epoch`):
zero_grad -> model(list of grids) -> loss.backward() -> [gradient all-reduce] -> clip_grad_norm_(0.1) -> AdamW.step()
-> OneCycleLR.step().  The data-parallel wrapper is explicit (one flat all-reduce) instead of torch DDP."""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch

from .optim import FusedAdamWClip, GradAllReducer


class MAEStepper:
    def __init__(self, model: torch.nn.Module, lr: float = 1e-4, weight_decay: float = 1e-3, clip_grad_norm: float = 0.1,
                 total_steps: Optional[int] = None, distributed: bool = False, process_group=None, n_buckets: int = 4):
        self.model = model
        params = [p for p in model.parameters() if p.requires_grad]
        self.optimizer = FusedAdamWClip(params, lr=lr, weight_decay=weight_decay, clip_grad_norm=clip_grad_norm)
        # run_swin_mae3d.py:594-598: OneCycleLR(max_lr=lr, total_steps=epochs*len(loader)); it also cycles beta1
        self.scheduler = None
        if total_steps is not None and total_steps > 1:
            self.scheduler = torch.optim.lr_scheduler.OneCycleLR(self.optimizer, max_lr=lr, total_steps=total_steps)
        self.reducer = GradAllReducer(params, process_group, n_buckets) if distributed else None

    def step(self, grids: Sequence[torch.Tensor], micro_batch: Optional[int] = None) -> torch.Tensor:
        """grids: list of (4,X,Y,Z) CUDA tensors.  Returns the device tensor [loss, loss_rgb, loss_alpha] (no sync).

        micro_batch: gradient accumulation - the list is processed in chunks of `micro_batch` grids whose gradients are averaged
        (exactly what data parallelism over len(grids)/micro_batch ranks computes: each rank's loss is normalised over its own
        grids, run_swin_mae3d.py:577-586,659-669), so a fixed global batch can be split over any number of GPUs (strong scaling).
        The gradient all-reduce is overlapped with the LAST micro-batch's backward."""
        grids = list(grids)
        mb = len(grids) if not micro_batch else int(micro_batch)
        chunks = [grids[i:i + mb] for i in range(0, len(grids), mb)]
        self.optimizer.zero_grad(set_to_none=True)
        total = None
        for ci, chunk in enumerate(chunks):
            loss, loss_rgb, loss_alpha = self.model(chunk)
            if self.reducer is not None and ci == len(chunks) - 1:
                self.reducer.arm()
            loss.backward()
            out = torch.stack([loss.detach(), loss_rgb.detach(), loss_alpha.detach()])
            total = out if total is None else total + out
        scale = 1.0 / len(chunks)
        if self.reducer is not None:
            flat = self.reducer.finish()
            self.optimizer.step(flat_grads=flat, flat_offsets=self.reducer.offsets, grad_scale=scale / self.reducer.world)
        else:
            self.optimizer.step(grad_scale=scale)
        if self.scheduler is not None:
            self.scheduler.step()
        return total * scale if len(chunks) > 1 else total

    def step_from_host(self, host_grids: List[torch.Tensor], device) -> List[float]:
        """End-to-end step: pinned host grids -> device copy -> step -> losses read back to the host."""
        dev = [g.to(device, non_blocking=True) for g in host_grids]      # run_swin_mae3d.py:656
        return self.step(dev).tolist()

    def steps_from_host(self, host_batches: Iterable[List[torch.Tensor]], device) -> List[List[float]]:
        """End-to-end steps over an iterable of batches of PINNED host grids (what the DataLoader of run_swin_mae3d.py:577-586
        yields with pin_memory): every batch is copied host->device and every step's loss triple is read back, but the copy of
        batch i+1 runs on a side stream while step i computes, and the losses of step i are fetched (pinned buffer + event)
        after step i+1 has been enqueued - the GPU never waits for the host."""
        dev = torch.device(device)
        main = torch.cuda.current_stream(dev)
        copy_stream = torch.cuda.Stream(dev)

        def upload(batch):
            # destination buffers come from the main stream's pool (no second pool, no cudaMalloc in the loop); the copy
            # stream first waits for the work already enqueued on the main stream, which may still be using those blocks
            tensors = [torch.empty(g.shape, dtype=g.dtype, device=dev) for g in batch]
            fence = torch.cuda.Event()
            fence.record(main)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(fence)
                for t, g in zip(tensors, batch):
                    t.copy_(g, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return tensors, ev

        # pinned read-back buffers are allocated once: cudaHostAlloc inside the loop waits for the GPU to drain
        host_bufs = [torch.empty(3, dtype=torch.float32, pin_memory=True) for _ in range(2)]
        n_done = 0
        results: List[List[float]] = []
        pending = None            # (pinned host tensor, event) of the previous step
        it = iter(host_batches)
        try:
            cur = upload(next(it))
        except StopIteration:
            return results
        while cur is not None:
            tensors, ev = cur
            main.wait_event(ev)
            try:
                cur = upload(next(it))          # overlaps with the step enqueued below
            except StopIteration:
                cur = None
            out = self.step(tensors)
            host_out = host_bufs[n_done & 1]
            n_done += 1
            host_out.copy_(out, non_blocking=True)
            done = torch.cuda.Event()
            done.record(main)
            if pending is not None:
                pending[1].synchronize()
                results.append(pending[0].tolist())
            pending = (host_out, done)
        pending[1].synchronize()
        results.append(pending[0].tolist())
        return results
