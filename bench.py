#!/usr/bin/env python
"""bench.py - grids/sec of the full MAE training step (fwd + loss + bwd + clip + AdamW [+ grad all-reduce]).

Workload (BASELINE.json metric): swin_s, synthetic 160^3 x 4 grids, 4 grids per GPU, mask_ratio 0.75, train mode
(stochastic depth on), fp32 storage.  One "step" = one optimiser step.  Default: weak scaling (every rank processes its own
`--batch` grids); `--global-batch G` fixes the grids per optimiser step over ALL ranks (strong scaling: each rank processes
G / world grids in micro-batches of `--batch` with gradient accumulation).  The only collective is the bucketed gradient
all-reduce, overlapped with the backward pass.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--model swin_s] [--res 160] [--batch 4]
                    [--global-batch G] [--precision fp16|bf16x3] [--workload train|fpn]

Besides the contract keys the line carries
  roofline            the dominant kernel (decoder1 3x3x3 convolution, forward launches) against the measured bf16 peak,
  wmsa                the tcgen05 W-MSA core launches of stage 1, timed inside the step,
  cpu_baseline        the reference algorithm on the host cores (oracle port), one grid,
  eager_gpu_baseline  the UNMODIFIED reference modules (baseline/_ref, staged by __graft_entry__.build) run as PyTorch eager on
                      the same GPU: torch-default TF32 convolutions and strict fp32 - the north star's "10x" denominator,
  parity_check        an eval forward of the timed batch against the live-reference known answer (tests/golden/kat_sized.json),
                      asserted at 1e-3 before anything is timed,
  peak_mem_gb         torch.cuda.max_memory_allocated over the timed region.

`--impl reference` times the reference algorithm on the host CPU (oracle port, all host threads) on a bounded sample of the same
workload: one grid per step.  `--workload fpn` is BASELINE config 5 (swin_l encoder + FPN feature extraction, inference).
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
METRIC_FPN = "grids/sec (160^3x4 swin_l encoder+FPN feature extraction, inference)"
FWD_GFLOP_PER_GRID = {"swin_s": 1377.3, "swin_t": 1322.5}   # BASELINE.md section 2 (160^3)
MODEL_TOL = 1e-3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="nmae", choices=["nmae", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "fpn"])
    ap.add_argument("--model", default=None, help="backbone (default swin_s for train, swin_l for fpn)")
    ap.add_argument("--res", type=int, default=160)
    ap.add_argument("--batch", type=int, default=None, help="grids per GPU and forward pass (default 4 train, 8 fpn)")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling: grids per optimiser step over all ranks (gradient accumulation in micro-batches of --batch)")
    ap.add_argument("--precision", default=None, choices=["bf16x3", "fp16"],
                    help="operand precision of the decoder's 3x3x3 convolutions (default: the library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    a = ap.parse_args()
    if a.model is None:
        a.model = "swin_s" if a.workload == "train" else "swin_l"
    if a.batch is None:
        a.batch = 4 if a.workload == "train" else 8
    return a


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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


def cpu_fpn_timer(model_name, res, steps, warmup, threads=None):
    """Config 5 on the host cores: oracle encoder + FPN forward of ONE grid (no_grad)."""
    import torch
    from oracle import nerf_mae_oracle as O
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.SWIN_CONFIGS[model_name]
    sd = O.init_state_dict(model_name, res, seed=0)
    C = cfg["embed_dim"]
    g = torch.Generator().manual_seed(0)
    neck = {}
    for i, c in enumerate([C, 2 * C, 4 * C, 8 * C]):
        neck[f"lateral_convs.{i}.weight"] = torch.randn(256, c, 1, 1, 1, generator=g) * 0.02
        neck[f"lateral_convs.{i}.bias"] = torch.zeros(256)
        neck[f"fpn_convs.{i}.weight"] = torch.randn(256, 256, 3, 3, 3, generator=g) * 0.01
        neck[f"fpn_convs.{i}.bias"] = torch.zeros(256)
    x = torch.rand(1, 4, res, res, res, generator=g)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.fpn_forward(neck, O.encoder_features(sd, x, cfg["depths"], cfg["num_heads"]))
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
    if args.workload == "fpn":
        val, cores, sec = cpu_fpn_timer(args.model, args.res, args.steps, args.warmup)
        metric, what = METRIC_FPN, f"{args.model} encoder + FPN forward"
    else:
        val, cores, sec = cpu_train_step_timer(args.model, args.res, args.steps, args.warmup)
        metric, what = METRIC, f"{args.model} MAE train step (fwd+bwd+clip+AdamW)"
    sample = f"1 grid/step ({what}, {args.res}^3), {args.steps} timed steps after {args.warmup} warm-up"
    line = {
        "impl": "reference", "metric": metric, "value": val, "unit": "grids/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong" if args.global_batch else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{what}, {args.res}^3x4 grids, mask_ratio 0.75, CPU oracle port of the reference",
                   "grids_per_step": 1, "torch": torch.__version__},
        "cpu_baseline": {"value": val, "unit": "grids/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "grids/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ eager GPU baseline
def import_reference():
    """The unmodified reference modules from baseline/_ref (staged by __graft_entry__.build()), or None."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "nerf_mae", "model", "mae")):
        return None
    import numpy
    if not hasattr(numpy, "float"):
        numpy.float = float                     # torch_utils.py:42 uses the alias numpy 2 removed
    if ref not in sys.path:
        sys.path.insert(0, ref)
    from nerf_mae.model.mae import swin_mae3d as R
    return R


# decoder1 3x3x3 convolution forward, DRAM bytes per launch from the committed ncu --set full capture (profiles/r2_ncu_full_conv3h.txt)
NCU_CONV_FWD_DRAM_BYTES = {("fp16", 160, 4, 48): 1.851546e9 + 3.094361e9}


def eager_gpu_baseline(model_name, res, B, steps=3, warmup=2):
    """The reference's own modules (SwinTransformer_MAE3D_New, unmodified) as PyTorch eager on cuda:0: the full train step of
    run_swin_mae3d.py:650-669 (zero_grad, forward, backward, clip_grad_norm_ 0.1, AdamW), twice: torch's default math
    (cudnn.allow_tf32 = True, matmul.allow_tf32 = False) and strict fp32 (both False)."""
    import torch
    import nerf_mae_b200 as N
    R = import_reference()
    if R is None:
        return {"value": None, "unit": "grids/s", "kind": "unavailable", "mode": "baseline/_ref is not staged (run __graft_entry__.build() "
                "where /root/reference exists)"}
    cfg = N.SWIN_CONFIGS[model_name]
    out = {"unit": "grids/s", "kind": "reference", "what": f"unmodified reference SwinTransformer_MAE3D_New ({model_name}), "
           f"{B} x {res}^3 grids, full train step, PyTorch eager on the same GPU", "steps": steps, "warmup": warmup}
    gen = torch.Generator().manual_seed(0)
    grids = [torch.rand(4, res, res, res, generator=gen).cuda() for _ in range(B)]
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for mode, (c_tf32, m_tf32) in (("tf32_default", (True, False)), ("strict_fp32", (False, False))):
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = c_tf32, m_tf32
            torch.manual_seed(0)
            random.seed(0)
            m = R.SwinTransformer_MAE3D_New([4, 4, 4], cfg["embed_dim"], cfg["depths"], cfg["num_heads"], [4, 4, 4], resolution=res,
                                            masking_prob=0.75).cuda().train()
            opt = torch.optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-3)
            torch.cuda.reset_peak_memory_stats()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for i in range(warmup + steps):
                if i == warmup:
                    torch.cuda.synchronize()
                    e0.record()
                opt.zero_grad()
                loss, _, _ = m(grids)
                loss.backward()
                torch.nn.utils.clip_grad_norm_(m.parameters(), 0.1)
                opt.step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[mode] = {"value": B / ms * 1e3, "ms_per_step": ms, "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9,
                         "loss_last": float(loss)}
            del m, opt, loss
            torch.cuda.empty_cache()
    except Exception as ex:
        out["error"] = f"{type(ex).__name__}: {ex}"
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    # headline of this leg = what a user of the reference gets out of the box
    out["mode"] = "tf32_default (cudnn.allow_tf32=True, matmul.allow_tf32=False: torch defaults)"
    out["value"] = out.get("tf32_default", {}).get("value")
    return out


# ------------------------------------------------------------------------------------------------ GPU arm
def parity_check(N, model, grids_dev):
    """Eval forward of the timed batch against the live-reference known answer; raises when off by more than 1e-3."""
    import torch
    with open(os.path.join(ROOT, "tests", "golden", "kat_sized.json")) as f:
        kat = json.load(f).get("bench_swin_s160_b4")
    if kat is None:
        return None
    model.eval()
    random.seed(42)
    with torch.no_grad():
        loss, lr, la, pred, valid, _ = model(grids_dev, is_eval=True)
    got = {"loss": float(loss), "loss_rgb": float(lr), "loss_alpha": float(la), "valid_sum": int(valid.sum()),
           "pred_sq_sum": float((pred.double() ** 2).sum())}
    del pred, valid
    model.train()
    rel = {k: abs(got[k] - kat[k]) / abs(kat[k]) for k in ("loss", "loss_rgb", "loss_alpha", "pred_sq_sum")}
    ok = all(v <= MODEL_TOL for v in rel.values()) and got["valid_sum"] == kat["valid_sum"]
    res = {"against": "live-reference KAT bench_swin_s160_b4 (oracle/make_golden_sized.py)", "tolerance": MODEL_TOL, "rel_err": rel,
           "valid_sum_exact": got["valid_sum"] == kat["valid_sum"], "loss": got["loss"], "loss_reference": kat["loss"], "ok": ok}
    if not ok:
        raise AssertionError(f"bench parity check failed: {res}")
    return res


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
        dist.init_process_group("nccl", device_id=dev)
    N.lib()
    N.set_conv_precision(args.precision)
    precision = N.get_conv_precision()
    if args.workload == "fpn":
        return run_fpn(args, N, _lib, world, rank, dev)

    B, R = args.batch, args.res
    strong = args.global_batch > 0
    if strong and args.global_batch % (world * 1) != 0:
        raise ValueError("--global-batch must be divisible by the number of GPUs")
    per_rank = args.global_batch // world if strong else B
    micro = min(B, per_rank)

    eager = None
    if world == 1 and rank == 0 and not args.no_eager:
        eager = eager_gpu_baseline(args.model, R, B)

    torch.manual_seed(0)                       # identical initial weights on every rank (what DDP's broadcast gives)
    model = N.build_model(args.model, R, 0.75).to(dev).train()
    total = (args.warmup + args.steps) * 2 + 16
    stepper = MAEStepper(model, lr=1e-4, weight_decay=1e-3, clip_grad_norm=0.1, total_steps=total, distributed=world > 1)
    gen = torch.Generator().manual_seed(rank)
    host = [torch.rand(4, R, R, R, generator=gen).pin_memory() for _ in range(per_rank)]
    grids = [h.to(dev) for h in host]
    h2d = sum(h.numel() * 4 for h in host)

    parity = None
    if rank == 0 and not args.no_parity and (args.model, R, per_rank) == ("swin_s", 160, 4):
        parity = parity_check(N, model, grids)
    torch.manual_seed(1000 + rank)             # per-rank stochastic-depth draws
    random.seed(rank)                          # per-rank mask draws

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

    def one_step():
        return stepper.step(grids, micro_batch=micro)

    for _ in range(args.warmup):
        out = one_step()
    sync()
    torch.cuda.reset_peak_memory_stats(dev)
    sampler = ClockSampler(local) if rank == 0 else None
    k0 = _lib.kernel_launches()
    _lib.timed_calls = {"nmae_conv3x3x3_fwd": [], "nmae_conv3x3x3_dgrad": [], "nmae_conv3x3x3_wgrad": [],
                        "nmae_conv3h_fwd": [], "nmae_conv3h_dgrad": [], "nmae_conv3h_wgrad": [],
                        "nmae_window_attention_fwd": [], "nmae_window_attention_bwd": []}
    ms = timed(one_step, args.steps)
    calls, _lib.timed_calls = _lib.timed_calls, None
    launches = _lib.kernel_launches() - k0
    clocks = sampler.stop() if sampler else None
    peak_mem = torch.cuda.max_memory_allocated(dev) / 1e9
    losses = out.tolist()
    grids_per_step = world * per_rank
    value = grids_per_step * args.steps / (ms / 1e3)

    e2e = None
    if not args.no_e2e:
        n_e2e = max(2, args.steps)
        if micro == per_rank:
            # untimed warm-up of the pipelined path itself (second set of device input buffers in the allocator pool, pinned read-back
            # buffers, the copy stream): those one-off costs were 13-76 ms depending on the box, i.e. up to 15 ms per step of a 5-step
            # measurement
            stepper.steps_from_host([host] * 2, dev)
            # every step: pinned host grids -> device, step, loss triple -> host; the copy of batch i+1 overlaps step i
            ms2 = timed(lambda: stepper.steps_from_host([host] * n_e2e, dev), 1)
        else:   # gradient accumulation: plain per-step upload + read-back
            stepper.step([h.to(dev, non_blocking=True) for h in host], micro_batch=micro).tolist()
            ms2 = timed(lambda: [stepper.step([h.to(dev, non_blocking=True) for h in host], micro_batch=micro).tolist()
                                 for _ in range(n_e2e)], 1)
        e2e = {"value": grids_per_step * n_e2e / (ms2 / 1e3), "unit": "grids/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
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
        sel = [e0.elapsed_time(e1) for e0, e1, ints in evs if ints[:6] == (micro, R, R, R, c1, c1)]
        if sel:
            dur[name] = sum(sel) / len(sel)
    conv_ms = sum(sum(e0.elapsed_time(e1) for e0, e1, _ in evs) for n, evs in calls.items() if "conv3" in n)
    # W-MSA core (tcgen05, csrc/wmsa_tc.cu): stage-1 launches (token grid (R/4)^3, embed_dim channels) timed inside the step.
    # useful FLOPs = QK^T + PV (+ the 5 products of the backward) over real 64x64x32 window-head blocks.
    wmsa = None
    H1, C1 = R // 4, model.embed_dim
    sel_f = [e0.elapsed_time(e1) for e0, e1, ints in calls["nmae_window_attention_fwd"] if ints[:5] == (micro, H1, H1, H1, C1)]
    sel_b = [e0.elapsed_time(e1) for e0, e1, ints in calls["nmae_window_attention_bwd"] if ints[:5] == (micro, H1, H1, H1, C1)]
    if sel_f and sel_b:
        nwin = ((H1 + 3) // 4) ** 3
        f_fwd = 2 * 2.0 * 64 * 64 * 32 * micro * nwin * (C1 // 32)
        wmsa = {"kernel": "wmsa_tc_fwd/bwd_kernel (stage 1: %d windows x %d heads x %d grids)" % (nwin, C1 // 32, micro),
                "fwd_ms": sum(sel_f) / len(sel_f), "bwd_ms": sum(sel_b) / len(sel_b),
                "fwd_useful_tflops": f_fwd / (sum(sel_f) / len(sel_f)) / 1e9, "bwd_useful_tflops": 2.5 * f_fwd / (sum(sel_b) / len(sel_b)) / 1e9,
                "all_stages_ms_per_step": sum(sum(e0.elapsed_time(e1) for e0, e1, _ in evs) for n, evs in calls.items()
                                              if "window_attention" in n) / args.steps,
                "tensor_pipe_pct": "not measured here (ncu metric): see the per-capture summaries under profiles/"}
    flops = 2.0 * micro * V * 27 * c1 * c1
    roof = None
    fwd_key = "nmae_conv3h_fwd" if precision == "fp16" else "nmae_conv3x3x3_fwd"
    if fwd_key in dur:
        ach = flops / (dur[fwd_key] * 1e-3) / 1e12
        kname = "conv3_h_kernel" if precision == "fp16" else "conv3_tc_kernel"
        roof = {"bound": "tensor", "kernel": f"{kname} (tcgen05 implicit-GEMM 3x3x3 conv, decoder1 {c1}->{c1} @{R}^3, fwd launches)",
                "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach / pk["tf_sust"],
                # dram__bytes_read.sum + dram__bytes_write.sum of one forward launch of exactly this shape, from the committed
                # ncu --set full capture (it cannot be measured inside a timed run); other shapes: null
                "traffic": NCU_CONV_FWD_DRAM_BYTES.get((precision, R, micro, c1)),
                "traffic_note": "ncu --set full capture of the same launch: profiles/r2_ncu_full_conv3h.txt (1.852 GB read + 3.094 GB written)",
                "traffic_algorithmic": micro * V * c1 * (2 if precision == "fp16" else 4) + micro * V * c1 * 4.0,
                "note": ("algorithmic FLOPs = 2*27*Cin*Cout per voxel; one fp16 pass per FLOP counted" if precision == "fp16" else
                         "fp32-equivalent FLOPs; the tensor pipe executes 3 bf16 passes per FLOP counted"),
                "peak_source": f"{pk['src']} bf16 sustained (MEASURED_PEAKS.json)",
                "ms_per_launch": dur, "flop_per_launch": flops,
                "share_of_step": conv_ms / ms}
    step_tflops = grids_per_step * 3 * FWD_GFLOP_PER_GRID.get(args.model, 0) * 1e-3 / (ms / args.steps / 1e3) if R == 160 else None

    cpu = None
    if world == 1 and not args.no_cpu_baseline and R > 160:
        cpu = {"value": None, "unit": "grids/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"skipped: one {R}^3 train step of the CPU port exceeds the bench's time budget (run --impl reference)"}
    elif world == 1 and not args.no_cpu_baseline:
        try:
            v, cores, sec = cpu_train_step_timer(args.model, R, 1, 0)
            cpu = {"value": v, "unit": "grids/s", "cores": cores, "kind": "port",
                   "sample": f"1 grid of the batch, 1 train step ({sec:.1f} s), no warm-up, oracle port of the reference on host cores"}
        except Exception as ex:  # the host baseline must not take the GPU number down with it
            cpu = {"value": None, "unit": "grids/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}

    line = {
        "metric": METRIC, "value": value, "unit": "grids/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.model} MAE train step (fwd+loss+bwd+clip0.1+AdamW), {R}^3x4 grids, {per_rank} grids/GPU/step"
                               f"{' in micro-batches of %d' % micro if micro != per_rank else ''}, mask_ratio 0.75, stochastic depth on, "
                               f"fp32 storage, 3x3x3 conv operands {precision}", "global_batch": grids_per_step,
                   "conv_precision": precision, "parallelism": f"dp{world}", "grad_allreduce": "4 buckets, overlapped with backward" if world > 1 else None,
                   "l2": "inputs_exceed_l2 (activations >> 126 MB; no flush needed)", "loss_last_warmup": losses},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "wmsa": wmsa, "cpu_baseline": cpu,
        "eager_gpu_baseline": eager, "parity_check": parity, "peak_mem_gb": peak_mem, "model_tflops_per_s": step_tflops,
    }
    if eager and eager.get("value"):
        line["vs_eager_gpu"] = {"value": value / eager["value"], "e2e": (e2e["value"] / eager["value"]) if e2e else None,
                                "vs_strict_fp32": value / eager["strict_fp32"]["value"] if eager.get("strict_fp32") else None}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_fpn(args, N, _lib, world, rank, dev):
    """BASELINE config 5: swin_l encoder + FPN(256) feature extraction, inference, `--batch` grids per GPU and step; replicas
    only (no collective).  e2e: the batch is copied from pinned host memory every step and a checksum of the coarsest level is
    read back."""
    import torch
    import torch.distributed as dist
    B, R = args.batch, args.res
    torch.manual_seed(0)
    m = N.SwinTransformer_FPN_Pretrained_Skip(resolution=R, is_eval=True, backbone_type=args.model).to(dev).eval()
    m.fpn_neck.init_weights()
    gen = torch.Generator().manual_seed(rank)
    host = torch.rand(B, 4, R, R, R, generator=gen).pin_memory()
    x = host.to(dev)

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
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    with torch.no_grad():
        for _ in range(args.warmup):
            outs = m(x)
        sync()
        torch.cuda.reset_peak_memory_stats(dev)
        sampler = ClockSampler(dev.index) if rank == 0 else None
        k0 = _lib.kernel_launches()
        ms = timed(lambda: m(x), args.steps)
        launches = _lib.kernel_launches() - k0
        clocks = sampler.stop() if sampler else None
        peak_mem = torch.cuda.max_memory_allocated(dev) / 1e9
        e2e = None
        if not args.no_e2e:
            def e2e_step():
                return float(m(host.to(dev, non_blocking=True))[-1].sum())
            e2e_step()
            ms2 = timed(e2e_step, args.steps)
            e2e = {"value": world * B * args.steps / (ms2 / 1e3), "unit": "grids/s", "h2d_bytes_per_step": host.numel() * 4,
                   "d2h_bytes_per_step": 4, "steps": args.steps}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            v, cores, sec = cpu_fpn_timer(args.model, R, 1, 0)
            cpu = {"value": v, "unit": "grids/s", "cores": cores, "kind": "port",
                   "sample": f"1 grid, 1 forward ({sec:.1f} s), oracle port of the reference encoder + FPN on host cores"}
        except Exception as ex:
            cpu = {"value": None, "unit": "grids/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
    line = {
        "metric": METRIC_FPN, "value": world * B * args.steps / (ms / 1e3), "unit": "grids/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.model} encoder + FPN(256) feature extraction (BASELINE config 5), {R}^3x4 grids, {B} grids/GPU/step, "
                               f"inference, fp32 storage, 3x3x3 conv operands {N.get_conv_precision()}", "global_batch": world * B,
                   "parallelism": f"replicas x{world} (no collective)", "l2": "inputs_exceed_l2", "outputs": [list(o.shape) for o in outs]},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "cpu_baseline": cpu, "peak_mem_gb": peak_mem,
        "roofline": None,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


class _OneLineStdout:
    """Rank 0 must print exactly ONE JSON line on stdout, but native libraries write there too (NCCL's version banner, ...):
    during the run file descriptor 1 points at stderr; print() is routed to the real stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)
        self.out = os.fdopen(self.real, "w", buffering=1)
        self.prev, sys.stdout = sys.stdout, self.out
        return self

    def __exit__(self, *exc):
        self.out.flush()
        sys.stdout = self.prev
        os.dup2(self.real, 1)
        return False


if __name__ == "__main__":
    a = parse()
    with _OneLineStdout():
        if a.impl == "reference":
            run_reference(a)
        else:
            run_nmae(a)
