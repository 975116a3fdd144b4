#!/usr/bin/env python
"""bench.py - grids/sec of the full MAE training step (fwd + loss + bwd + clip + AdamW [+ grad all-reduce]).

Workload (BASELINE.json metric): swin_s, synthetic 160^3 x 4 grids, 4 grids per GPU, mask_ratio 0.75, train mode
(stochastic depth on), fp32.  One "step" = one optimiser step over the per-GPU batch.  Weak scaling: every rank
processes its own 4 grids; the only collective is the flat gradient all-reduce.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--model swin_s] [--res 160] [--batch 4]

Besides the contract keys the line carries `roofline` (the dominant kernel, conv3_tc_kernel, against the measured bf16 peak),
`wmsa` (the tcgen05 W-MSA core launches of stage 1, timed inside the step) and `cpu_baseline`.

`--impl reference` times the reference algorithm on the host CPU (the oracle port of the reference's PyTorch path,
all host threads) on a bounded sample of the same workload: one grid per step.
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "grids/sec (160^3x4 swin_s MAE train step)"
FWD_GFLOP_PER_GRID = {"swin_s": 1377.3, "swin_t": 1322.5}   # BASELINE.md section 2 (160^3)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="nmae", choices=["nmae", "reference"])
    ap.add_argument("--model", default="swin_s")
    ap.add_argument("--res", type=int, default=160)
    ap.add_argument("--batch", type=int, default=4, help="grids per GPU")
    ap.add_argument("--precision", default=None, choices=["bf16x3", "fp16"],
                    help="operand precision of the decoder's 3x3x3 convolutions (default: the library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_train_step_timer(model_name, res, steps, warmup, threads=None):
    """The reference's algorithm (oracle port of its PyTorch path) on the host cores: fwd+bwd+clip+AdamW for ONE grid per
    step.  Returns (grids_per_sec, cores, seconds_per_step)."""
    import torch
    from oracle import nerf_mae_oracle as O
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.SWIN_CONFIGS[model_name]
    sd = O.init_state_dict(model_name, res, seed=0)
    for k, v in sd.items():
        v.requires_grad_(v.dtype.is_floating_point and k != "pos_embed")
    g = torch.Generator().manual_seed(0)
    grid = torch.rand(4, res, res, res, generator=g)
    state = {}
    random.seed(0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.train_step(sd, state, [grid], cfg["depths"], cfg["num_heads"], res, 0.75, lr=1e-4)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return 1.0 / sec, cores, sec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    val, cores, sec = cpu_train_step_timer(args.model, args.res, args.steps, args.warmup)
    sample = f"1 grid/step ({args.model} {args.res}^3, fwd+bwd+clip+AdamW), {args.steps} timed steps after {args.warmup} warm-up"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "grids/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.model} MAE train step, {args.res}^3x4 grids, mask_ratio 0.75, CPU oracle port of the reference",
                   "grids_per_step": 1, "torch": torch.__version__},
        "cpu_baseline": {"value": val, "unit": "grids/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "grids/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_nmae(args):
    import torch
    import torch.distributed as dist

    import nerf_mae_b200 as N
    from nerf_mae_b200 import _lib
    from nerf_mae_b200.trainer import MAEStepper

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: nerf-mae_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")   # keeps NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    N.lib()
    N.set_conv_precision(args.precision)
    precision = N.get_conv_precision()
    hk = "h" if precision == "fp16" else "x3x3"      # nmae_conv3h_* / nmae_conv3x3x3_* entry points

    B, R = args.batch, args.res
    torch.manual_seed(0)                       # identical initial weights on every rank (what DDP's broadcast gives)
    model = N.build_model(args.model, R, 0.75).to(dev).train()
    total = args.warmup * 2 + args.steps * 2 + 8
    stepper = MAEStepper(model, lr=1e-4, weight_decay=1e-3, clip_grad_norm=0.1, total_steps=total, distributed=world > 1)
    torch.manual_seed(1000 + rank)             # per-rank data and stochastic-depth draws
    random.seed(rank)                          # per-rank mask draws
    gen = torch.Generator().manual_seed(rank)
    host = [torch.rand(4, R, R, R, generator=gen).pin_memory() for _ in range(B)]
    grids = [h.to(dev) for h in host]
    h2d = sum(h.numel() * 4 for h in host)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, n):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(args.warmup):
        out = stepper.step(grids)
    sync()
    sampler = ClockSampler(local) if rank == 0 else None
    k0 = _lib.kernel_launches()
    _lib.timed_calls = {"nmae_conv3x3x3_fwd": [], "nmae_conv3x3x3_dgrad": [], "nmae_conv3x3x3_wgrad": [],
                        "nmae_conv3h_fwd": [], "nmae_conv3h_dgrad": [], "nmae_conv3h_wgrad": [],
                        "nmae_window_attention_fwd": [], "nmae_window_attention_bwd": []}
    ms = timed(lambda: stepper.step(grids), args.steps)
    calls, _lib.timed_calls = _lib.timed_calls, None
    launches = _lib.kernel_launches() - k0
    clocks = sampler.stop() if sampler else None
    losses = out.tolist()
    value = world * B * args.steps / (ms / 1e3)

    e2e = None
    if not args.no_e2e:
        n_e2e = max(2, args.steps)
        stepper.step_from_host(host, dev)
        # every step: pinned host grids -> device, step, loss triple -> host; the copy of batch i+1 overlaps step i
        ms2 = timed(lambda: stepper.steps_from_host([host] * n_e2e, dev), 1)
        e2e = {"value": world * B * n_e2e / (ms2 / 1e3), "unit": "grids/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
               "steps": n_e2e}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    # dominant kernel: the full-resolution 3x3x3 convolution of decoder1 (implicit GEMM), forward launches
    V = R ** 3
    c1 = model.embed_dim // 2
    dur = {}
    for name, evs in calls.items():
        if "conv3" not in name:
            continue
        sel = [e0.elapsed_time(e1) for e0, e1, ints in evs if ints[:6] == (B, R, R, R, c1, c1)]
        if sel:
            dur[name] = sum(sel) / len(sel)
    conv_ms = sum(sum(e0.elapsed_time(e1) for e0, e1, _ in evs) for n, evs in calls.items() if "conv3" in n)
    # W-MSA core (tcgen05, csrc/wmsa_tc.cu): stage-1 launches (token grid (R/4)^3, embed_dim channels) timed inside the step.
    # useful FLOPs = QK^T + PV (+ the 5 products of the backward) over real 64x64x32 window-head blocks.
    wmsa = None
    H1, C1 = R // 4, model.embed_dim
    sel_f = [e0.elapsed_time(e1) for e0, e1, ints in calls["nmae_window_attention_fwd"] if ints[:5] == (B, H1, H1, H1, C1)]
    sel_b = [e0.elapsed_time(e1) for e0, e1, ints in calls["nmae_window_attention_bwd"] if ints[:5] == (B, H1, H1, H1, C1)]
    if sel_f and sel_b:
        nwin = ((H1 + 3) // 4) ** 3
        f_fwd = 2 * 2.0 * 64 * 64 * 32 * B * nwin * (C1 // 32)
        wmsa = {"kernel": "wmsa_tc_fwd/bwd_kernel (stage 1: %d windows x %d heads x %d grids)" % (nwin, C1 // 32, B),
                "fwd_ms": sum(sel_f) / len(sel_f), "bwd_ms": sum(sel_b) / len(sel_b),
                "fwd_useful_tflops": f_fwd / (sum(sel_f) / len(sel_f)) / 1e9, "bwd_useful_tflops": 2.5 * f_fwd / (sum(sel_b) / len(sel_b)) / 1e9,
                "all_stages_ms_per_step": sum(sum(e0.elapsed_time(e1) for e0, e1, _ in evs) for n, evs in calls.items()
                                              if "window_attention" in n) / args.steps,
                "tensor_pipe_pct_ncu": {"fwd": 9.3, "bwd": 8.2, "source": "profiles/r1_ncu_full_wmsa.txt (sm__pipe_tensor_cycles_active.avg."
                                        "pct_of_peak_sustained_elapsed; captured separately, never timed under the profiler)"}}
    flops = 2.0 * B * V * 27 * c1 * c1
    roof = None
    fwd_key = "nmae_conv3h_fwd" if precision == "fp16" else "nmae_conv3x3x3_fwd"
    if fwd_key in dur:
        ach = flops / (dur[fwd_key] * 1e-3) / 1e12
        # DRAM bytes per launch of this kernel from the committed `ncu --set full` capture (dram__bytes_read+write.sum:
        # 4.38 GB bf16 hi/lo operand image read + 3.12 GB fp32 output written), valid for the captured shape only
        # (B=4, 160^3, 48->48): profiles/r1_ncu_full_conv3.txt
        traffic = 7.494e9 if (B, R, c1) == (4, 160, 48) and precision == "bf16x3" else None
        kname = "conv3_h_kernel" if precision == "fp16" else "conv3_tc_kernel"
        roof = {"bound": "tensor", "kernel": f"{kname} (tcgen05 implicit-GEMM 3x3x3 conv, decoder1 {c1}->{c1} @{R}^3, fwd launches)",
                "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach / pk["tf_sust"], "traffic": traffic,
                "traffic_algorithmic": 2.0 * B * V * c1 * 4,
                "note": ("single fp16 pass per FLOP counted" if precision == "fp16" else
                         "fp32-equivalent FLOPs; the tensor pipe executes 3 bf16 passes per FLOP counted"),
                "peak_source": f"{pk['src']} bf16 sustained (MEASURED_PEAKS.json)",
                "ms_per_launch": dur, "flop_per_launch": flops,
                "share_of_step": conv_ms / ms}
    step_tflops = world * B * 3 * FWD_GFLOP_PER_GRID.get(args.model, 0) * 1e-3 / (ms / args.steps / 1e3) if R == 160 else None

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            v, cores, sec = cpu_train_step_timer(args.model, R, 1, 0)
            cpu = {"value": v, "unit": "grids/s", "cores": cores, "kind": "port",
                   "sample": f"1 grid of the batch, 1 train step ({sec:.1f} s), no warm-up, oracle port of the reference on host cores"}
        except Exception as ex:  # the host baseline must not take the GPU number down with it
            cpu = {"value": None, "unit": "grids/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}

    line = {
        "metric": METRIC, "value": value, "unit": "grids/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{args.model} MAE train step (fwd+loss+bwd+clip0.1+AdamW), {R}^3x4 grids, {B} grids/GPU, "
                               f"mask_ratio 0.75, stochastic depth on, fp32 storage, 3x3x3 conv operands {precision}", "global_batch": world * B,
                   "conv_precision": precision,
                   "parallelism": f"dp{world}", "l2": "inputs_exceed_l2 (activations >> 126 MB; no flush needed)",
                   "loss_last_warmup": losses},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "wmsa": wmsa, "cpu_baseline": cpu,
        "model_tflops_per_s": step_tflops,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_nmae(a)
