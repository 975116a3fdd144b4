"""Pretraining / evaluation driver with the reference's CLI and Trainer surface (nerf_mae/run_swin_mae3d.py), running the
B200-native model.  Only the MAE-relevant flags are kept (the reference inherits ~60 NeRF-RPN flags that the MAE path
never reads); behaviour differences are deliberate and listed here:

  * data parallelism is explicit: one process per GPU (torchrun or --gpus with mp.spawn), identical initial weights from
    a common seed, ONE flat fp32 gradient all-reduce per step (optim.GradAllReducer) instead of torch DDP buckets, no
    per-step barrier, logging scalars reduced only every --log_interval steps;
  * clip_grad_norm_(0.1) + AdamW run as one fused multi-tensor kernel pair (optim.FusedAdamWClip); OneCycleLR is torch's;
  * --backbone_type is honoured (the reference hard-codes swin_s, run_swin_mae3d.py:377);
  * validation runs on rank 0 through the bare module (the reference calls the DDP wrapper from one rank only);
  * checkpoints keep the reference format {"epoch","state_dict","train_args"} and additionally hold optimizer /
    scheduler / RNG state under "resume" so that --checkpoint really resumes;
  * --dataset synthetic generates torch.rand grids (no dataset ships in this environment); npz scenes are read exactly
    like nerf_rpn/datasets.py:52-108 (rgbsigma (W,L,H,4) float or uint8, density -> alpha, transpose to (4,W,L,H)).
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import random
import sys
import time

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
import nerf_mae_b200 as N  # noqa: E402  (import shim at the repository root)
from nerf_mae_b200.optim import FusedAdamWClip, GradAllReducer  # noqa: E402


# reference flags that only the detection (NeRF-RPN / FCOS) heads read: accepted so that the reference's command lines run
# unchanged, never used by the MAE path (nerf_mae/run_swin_mae3d.py:41-313 inherits them from run_fcos.py)
_UNUSED_REFERENCE_FLAGS = {
    "boxes_path": None, "train_csv": "", "val_csv": "", "test_csv": "", "mae_checkpoint": "", "ap_top_n": None,
    "center_sampling_radius": 1.5, "filter": "none", "filter_threshold": 0.7, "fpn_post_nms_top_n": 2500, "input_dim": 4,
    "iou_loss_type": "iou", "min_size": 0.0, "nms_thresh": 0.3, "num_convs": 4, "pre_nms_thresh": 0.0, "pre_nms_top_n": 2500,
    "proj2d_loss_weight": 0.0, "reg_loss_weight": 1.0, "rot_scale_prob": 0.5,
}
_UNUSED_REFERENCE_SWITCHES = ["centerness_on_reg", "conv_at_start", "load_backbone_only", "norm_reg_targets", "output_all",
                              "output_proposals", "output_voxel_scores", "rotated_bbox", "save_level_index", "train_all",
                              "use_additional_l1_loss", "preload"]


def parse_args(argv=None):
    """The reference CLI (nerf_mae/run_swin_mae3d.py:41-313) with the reference's defaults; flags the MAE path never reads are
    accepted and ignored (see _UNUSED_REFERENCE_FLAGS).  Extensions: --dataset synthetic, --synthetic_scenes, --gpu_ingest,
    --seed, --conv_precision."""
    p = argparse.ArgumentParser(description="B200-native NeRF-MAE pretraining (reference CLI: nerf_mae/run_swin_mae3d.py:41-313)")
    p.add_argument("--mode", default="train", choices=["train", "eval", "benchmark"])
    p.add_argument("--dataset", "--dataset_name", default="hypersim",
                   choices=["hypersim", "front3d", "general", "scannet", "hm3d", "synthetic"])
    p.add_argument("--features_path", default="")
    p.add_argument("--dataset_split", default="")
    p.add_argument("--save_path", default="")
    p.add_argument("--checkpoint", default=None)
    p.add_argument("--backbone_type", default="swin_s", choices=list(N.SWIN_CONFIGS))
    p.add_argument("--masking_prob", default=0.5, type=float)
    p.add_argument("--masking_strategy", default="random")
    p.add_argument("--resolution", default=160, type=int)
    p.add_argument("--normalize_density", action="store_true")
    p.add_argument("--batch_size", default=1, type=int, help="GLOBAL batch size (split over the ranks, as in the reference)")
    p.add_argument("--num_epochs", default=100, type=int)
    p.add_argument("--lr", default=5e-3, type=float)
    p.add_argument("--weight_decay", default=0.01, type=float)
    p.add_argument("--clip_grad_norm", default=0.1, type=float)
    p.add_argument("--log_interval", default=20, type=int)
    p.add_argument("--eval_interval", default=1, type=int)
    p.add_argument("--keep_checkpoints", default=1, type=int, help="epoch_*.pt files kept besides model_best.pt")
    p.add_argument("--gpus", default="")
    p.add_argument("--percent_train", default=1.0, type=float)
    p.add_argument("--flip_prob", default=0.5, type=float)
    p.add_argument("--rotate_prob", default=0.5, type=float)
    p.add_argument("--log_to_file", action="store_true", help="also write the log to <save_path>/train.log")
    p.add_argument("--wandb", action="store_true")
    p.add_argument("--tags", default="")
    for name, default in _UNUSED_REFERENCE_FLAGS.items():
        p.add_argument("--" + name, default=default, type=type(default) if default is not None else str, help=argparse.SUPPRESS)
    for name in _UNUSED_REFERENCE_SWITCHES:
        p.add_argument("--" + name, action="store_true", help=argparse.SUPPRESS)
    # extensions
    p.add_argument("--gpu_ingest", action="store_true",
                   help="ship the raw (W,L,H,4) arrays to the GPU and decode / augment / pad there (nmae_ingest_scene) "
                        "instead of on the loader's CPU workers")
    p.add_argument("--synthetic_scenes", default=32, type=int)
    p.add_argument("--seed", default=0, type=int)
    p.add_argument("--conv_precision", default=None, choices=list(N.functional.CONV_PRECISIONS),
                   help="operand precision of the decoder's 3x3x3 convolutions (default: the library default)")
    return p.parse_args(argv)


# ------------------------------------------------------------------------------------------------ data
def density_to_alpha(density):
    return np.clip(1.0 - np.exp(-np.exp(density) / 100.0), 0.0, 1.0)      # datasets.py:246-248


def load_scene_features(path, normalize_density=True):
    """One scene file -> (4, W, L, H) tensor, exactly as the reference loads it (nerf_rpn/datasets.py:88-104): `rgbsigma`
    (W, L, H, 4), density -> alpha on the last channel written back INTO the stored array (so its dtype decides the rounding),
    channels first, uint8 grids scaled by 1/255."""
    with np.load(path) as f:
        rgbsigma = f["rgbsigma"]
        if normalize_density:
            rgbsigma[..., -1] = density_to_alpha(rgbsigma[..., -1])
        t = torch.from_numpy(np.transpose(rgbsigma, (3, 0, 1, 2)))
        if t.dtype == torch.uint8:
            t = t.float() / 255.0
    return t


def draw_augmentation(flip_prob, rotate_prob):
    """The three RNG draws of augment_grid without touching any data: (rotate, flip_axis1, flip_axis2) for
    functional.ingest_scenes(), which applies them as an index map on the GPU."""
    rot = random.random() < rotate_prob
    return rot, random.random() < flip_prob, random.random() < flip_prob


def augment_grid(t, flip_prob, rotate_prob):
    """The reference's scene augmentation for box-free, z-up grids (nerf_rpn/datasets.py:172-234 with boxes=None): one draw
    for the 90-degree rotation in the (W, L) plane (transpose then flip W), then one draw per horizontal axis for the flips -
    three `random.random()` draws per scene, in this order, from the global Mersenne-Twister stream."""
    if not 0 <= flip_prob <= 1:
        raise ValueError("flip_prob must be between 0 and 1, but got {}".format(flip_prob))
    if not 0 <= rotate_prob <= 1:
        raise ValueError("rotate_prob must be between 0 and 1, but got {}".format(rotate_prob))
    if random.random() < rotate_prob:
        t = torch.flip(torch.transpose(t, 1, 2), [1])
    for axis in (1, 2):
        if random.random() < flip_prob:
            t = t.flip(dims=[axis])
    return t


class SceneDataset(torch.utils.data.Dataset):
    """npz scenes (datasets.py:52-108) with the flip / rot90 augmentation (datasets.py:172-234), or synthetic grids."""

    def __init__(self, args, scenes, train: bool):
        self.args, self.scenes, self.train = args, scenes, train

    def __len__(self):
        return len(self.scenes)

    def __getitem__(self, i):
        a = self.args
        if a.dataset == "synthetic":
            g = torch.Generator().manual_seed(1000003 * a.seed + int(self.scenes[i]))
            ext = [a.resolution - int(torch.randint(0, a.resolution // 4 + 1, (1,), generator=g)) for _ in range(3)]
            return torch.rand(4, *ext, generator=g), None, str(self.scenes[i])
        path = os.path.join(a.features_path, self.scenes[i] + ".npz")
        if getattr(a, "gpu_ingest", False):
            # raw array as stored + the augmentation decisions (same RNG draws, same order); the arithmetic happens on the GPU
            with np.load(path) as f:
                raw = torch.from_numpy(np.ascontiguousarray(f["rgbsigma"]))
            flags = draw_augmentation(a.flip_prob, a.rotate_prob) if self.train and (a.flip_prob > 0 or a.rotate_prob > 0) \
                else (False, False, False)
            return raw, flags, self.scenes[i]
        t = load_scene_features(path, a.normalize_density)
        if self.train and (a.flip_prob > 0 or a.rotate_prob > 0):
            t = augment_grid(t, a.flip_prob, a.rotate_prob)
        return t.contiguous(), None, self.scenes[i]

    @staticmethod
    def collate_fn(batch):
        return [b[0] for b in batch], [b[1] for b in batch], [b[2] for b in batch]


def scene_lists(args):
    if args.dataset == "synthetic":
        n = args.synthetic_scenes
        ids = list(range(n))
        return ids[: max(1, int(0.8 * n))], ids[max(1, int(0.8 * n)):] or ids[:1]
    with np.load(args.dataset_split) as split:                                   # run_swin_mae3d.py:413-469
        train, val = list(split["train_scenes"]), list(split["val_scenes"])
    if args.percent_train < 1.0:
        train = train[: max(1, int(len(train) * args.percent_train))]
    return train, val


def mse(pred, target, valid_mask=None):                                            # nerf_rpn/model/metrics.py:69-76
    d = (pred - target) ** 2
    if valid_mask is not None:
        d = d[valid_mask.expand_as(d)]
    return d.mean()


def psnr(pred, target, valid_mask=None):                                           # nerf_rpn/model/metrics.py:78-79
    return -10.0 * torch.log10(mse(pred, target, valid_mask))


# ------------------------------------------------------------------------------------------------ trainer
class Trainer:
    def __init__(self, args, rank=0, world_size=1, device_id=0):
        self.args, self.rank, self.world_size = args, rank, world_size
        self.device = torch.device("cuda", device_id)
        self.logger = logging.getLogger(f"worker_{rank}")
        if getattr(args, "conv_precision", None):
            N.set_conv_precision(args.conv_precision)
        if getattr(args, "log_to_file", False) and rank == 0 and args.save_path:
            os.makedirs(args.save_path, exist_ok=True)
            self.logger.addHandler(logging.FileHandler(os.path.join(args.save_path, "train.log")))
        self.wandb = None
        if getattr(args, "wandb", False) and rank == 0:
            try:
                import wandb
                self.wandb = wandb
                wandb.init(project="nerf-mae", tags=[t for t in args.tags.split(",") if t], config=vars(args))
            except Exception as ex:                      # not installed / offline: say so instead of silently dropping --wandb
                self.logger.warning(f"--wandb requested but unavailable ({ex}); logging to the console only")
        ignored = [k for k, d in _UNUSED_REFERENCE_FLAGS.items() if getattr(args, k, d) != d] + \
                  [k for k in _UNUSED_REFERENCE_SWITCHES if getattr(args, k, False)]
        if ignored and rank == 0:
            self.logger.warning("flags not used by the MAE path are ignored: " + ", ".join(ignored))
        torch.manual_seed(args.seed)                 # identical initial weights on every rank (DDP's rank-0 broadcast)
        self.model = self.build_model()
        self.start_epoch = 0
        self._resume = None
        if args.checkpoint:
            assert os.path.exists(args.checkpoint), "The checkpoint does not exist."     # run_swin_mae3d.py:338
            ck = torch.load(args.checkpoint, map_location="cpu", weights_only=False)
            self.model.load_state_dict(ck["state_dict"])
            self._resume = ck.get("resume")
            self.start_epoch = int(ck.get("epoch", -1)) + 1 if self._resume else 0
        self.model.to(self.device)
        torch.manual_seed(args.seed + 7919 * (rank + 1))   # per-rank stochastic-depth stream
        random.seed(args.seed + rank)                      # per-rank mask stream (each rank draws its own, as in the reference)
        self.best = -1e9
        self.history = []                                  # (epoch, step, loss, loss_rgb, loss_alpha) at every log point (rank 0)
        if self._resume:                                   # true resume: the RNG streams continue where the checkpoint left them
            r = self._resume
            if r.get("py_rng") is not None:
                random.setstate(r["py_rng"])
            if r.get("torch_rng") is not None:
                torch.set_rng_state(r["torch_rng"])
            if r.get("cuda_rng") is not None:
                torch.cuda.set_rng_state(r["cuda_rng"], self.device)
            self.best = float(r.get("best", -1e9))

    def build_model(self):
        a = self.args
        return N.build_model(a.backbone_type, a.resolution, a.masking_prob, masking_strategy=a.masking_strategy)

    # ---- checkpoints (run_swin_mae3d.py:471-489)
    def save_checkpoint(self, epoch, path, optimizer=None, scheduler=None):
        if self.rank != 0:
            return
        ck = {"epoch": epoch, "state_dict": self.model.state_dict(), "train_args": vars(self.args)}
        if optimizer is not None:
            ck["resume"] = {"optimizer": optimizer.state_dict(), "scheduler": scheduler.state_dict() if scheduler else None,
                            "steps": optimizer._steps, "torch_rng": torch.get_rng_state(), "py_rng": random.getstate(),
                            "cuda_rng": torch.cuda.get_rng_state(self.device), "best": self.best}
        torch.save(ck, path)

    def _prune_checkpoints(self):
        """--keep_checkpoints (run_swin_mae3d.py:489-499): keep the newest N epoch_*.pt files."""
        keep = max(0, int(getattr(self.args, "keep_checkpoints", 1)))
        files = [f for f in os.listdir(self.args.save_path) if f.startswith("epoch_") and f.endswith(".pt")]
        files.sort(key=lambda f: int(f[len("epoch_"):-3]))
        for f in files[:max(0, len(files) - keep)]:
            os.remove(os.path.join(self.args.save_path, f))

    def train_loop(self):
        a = self.args
        train_scenes, val_scenes = scene_lists(a)
        per_rank = max(1, a.batch_size // self.world_size)                             # run_swin_mae3d.py:577-586
        train_set = SceneDataset(a, train_scenes, True)
        sampler = torch.utils.data.distributed.DistributedSampler(train_set, self.world_size, self.rank, shuffle=True) \
            if self.world_size > 1 else None
        loader = torch.utils.data.DataLoader(train_set, batch_size=per_rank, shuffle=sampler is None, sampler=sampler,
                                             collate_fn=SceneDataset.collate_fn, num_workers=2, pin_memory=True, drop_last=False)
        params = [p for p in self.model.parameters() if p.requires_grad]
        self.optimizer = FusedAdamWClip(params, lr=a.lr, weight_decay=a.weight_decay, clip_grad_norm=a.clip_grad_norm)
        total = max(2, a.num_epochs * len(loader))
        self.scheduler = torch.optim.lr_scheduler.OneCycleLR(self.optimizer, max_lr=a.lr, total_steps=total)   # :594-598
        self.reducer = GradAllReducer(params) if self.world_size > 1 else None
        if self._resume:
            self.optimizer.load_state_dict(self._resume["optimizer"])       # restores the moments and the step count
            # OneCycleLR bakes total_steps into its state: fast-forward a fresh schedule instead of loading the old one,
            # so that a run resumed with a different --num_epochs keeps a valid schedule
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                for _ in range(min(int(self._resume["steps"]), total - 1)):
                    self.scheduler.step()
        os.makedirs(a.save_path or ".", exist_ok=True)
        for epoch in range(self.start_epoch, a.num_epochs):
            if sampler is not None:
                sampler.set_epoch(epoch)
            self.train_epoch(epoch, loader)
            if (epoch + 1) % a.eval_interval == 0 or epoch == a.num_epochs - 1:
                if self.rank == 0:
                    m = self.eval(SceneDataset(a, val_scenes, False))
                    self.logger.info(f"epoch {epoch}: val psnr {m['psnr']:.3f} mse {m['mse']:.6f} loss {m['loss']:.5f}")
                    if self.wandb is not None:
                        self.wandb.log({"val/" + k: v for k, v in m.items()}, step=epoch)
                    if m["psnr"] > self.best:
                        self.best = m["psnr"]
                        self.save_checkpoint(epoch, os.path.join(a.save_path, "model_best.pt"), self.optimizer, self.scheduler)
                    self.save_checkpoint(epoch, os.path.join(a.save_path, f"epoch_{epoch}.pt"), self.optimizer, self.scheduler)
                    self._prune_checkpoints()
                if self.world_size > 1:
                    dist.barrier()

    def train_epoch(self, epoch, loader):                                             # run_swin_mae3d.py:644-709
        a = self.args
        self.model.train()
        t0, seen = time.time(), 0
        for step, (rgbsigma, flags, _) in enumerate(loader):
            grids = [g.to(self.device, non_blocking=True) for g in rgbsigma]
            self.optimizer.zero_grad(set_to_none=True)
            if getattr(a, "gpu_ingest", False) and a.dataset != "synthetic":
                xb, ext = N.functional.ingest_scenes(grids, a.resolution, a.normalize_density, flags)
                loss, loss_rgb, loss_alpha = self.model.forward_padded(xb, ext)
            else:
                loss, loss_rgb, loss_alpha = self.model(grids)
            if self.reducer is not None:
                self.reducer.arm()                   # buckets are all-reduced while the rest of the backward still runs
            loss.backward()
            if self.reducer is not None:
                flat = self.reducer.finish()
                self.optimizer.step(flat_grads=flat, flat_offsets=self.reducer.offsets, grad_scale=1.0 / self.world_size)
            else:
                self.optimizer.step()
            self.scheduler.step()
            seen += len(grids)
            if (step + 1) % a.log_interval == 0 or step == len(loader) - 1:
                stats = torch.stack([loss.detach(), loss_rgb.detach(), loss_alpha.detach()])
                if self.world_size > 1:
                    dist.all_reduce(stats)
                    stats /= self.world_size
                if self.rank == 0:
                    l, lr_, la = stats.tolist()
                    self.history.append((epoch, step, l, lr_, la))
                    if self.wandb is not None:
                        self.wandb.log({"train/loss": l, "train/loss_rgb": lr_, "train/loss_alpha": la})
                    self.logger.info(f"epoch {epoch} step {step + 1}/{len(loader)} lr {self.scheduler.get_last_lr()[0]:.3e} "
                                     f"loss {l:.5f} rgb {lr_:.5f} alpha {la:.5f} | {self.world_size * seen / (time.time() - t0):.2f} grids/s")

    @torch.no_grad()
    def eval(self, dataset):                                                          # run_swin_mae3d.py:711-806
        self.model.eval()
        loader = torch.utils.data.DataLoader(dataset, batch_size=max(1, self.args.batch_size // self.world_size), shuffle=False,
                                             collate_fn=SceneDataset.collate_fn, num_workers=2)
        tot = {"psnr": 0.0, "mse": 0.0, "loss": 0.0}
        n = 0
        for rgbsigma, _, _ in loader:
            grids = [g.to(self.device) for g in rgbsigma]
            loss, loss_rgb, loss_alpha, pred, _, target = self.model(grids, is_eval=True)
            mask = target[..., 3:4] > 0.01
            tot["psnr"] += float(psnr(pred[..., :3], target[..., :3], mask))
            tot["mse"] += float(mse(pred[..., :3], target[..., :3], mask))
            tot["loss"] += float(loss)
            n += 1
        out = {k: v / max(n, 1) for k, v in tot.items()}
        if self.args.mode == "eval":
            os.makedirs(self.args.save_path, exist_ok=True)
            with open(os.path.join(self.args.save_path, "eval.json"), "w") as f:
                json.dump(out, f)
        return out

    def benchmark(self, steps=10, warmup=3):
        """--mode benchmark is accepted but never implemented by the reference (run_swin_mae3d.py:47,844-847)."""
        a = self.args
        per_rank = max(1, a.batch_size // self.world_size)
        grids = [torch.rand(4, a.resolution, a.resolution, a.resolution, device=self.device) for _ in range(per_rank)]
        params = [p for p in self.model.parameters() if p.requires_grad]
        opt = FusedAdamWClip(params, lr=a.lr, weight_decay=a.weight_decay, clip_grad_norm=a.clip_grad_norm)
        red = GradAllReducer(params) if self.world_size > 1 else None
        self.model.train()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for i in range(warmup + steps):
            if i == warmup:
                torch.cuda.synchronize()
                ev[0].record()
            opt.zero_grad(set_to_none=True)
            loss, _, _ = self.model(grids)
            loss.backward()
            if red is not None:
                opt.step(flat_grads=red.reduce(), flat_offsets=red.offsets, grad_scale=1.0 / self.world_size)
            else:
                opt.step()
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / steps
        if self.rank == 0:
            self.logger.info(f"benchmark: {ms:.1f} ms/step, {self.world_size * per_rank / ms * 1e3:.2f} grids/s")
        return ms


def main_worker(rank, world_size, gpu_ids, args, port):
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(name)s %(message)s")
    if world_size > 1:
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world_size)
    dev = gpu_ids[rank]
    torch.cuda.set_device(dev)
    tr = Trainer(args, rank, world_size, dev)
    if args.mode == "train":
        tr.train_loop()
    elif args.mode == "eval":
        if rank == 0:
            _, val = scene_lists(args)
            print(json.dumps(tr.eval(SceneDataset(args, val, False))))
    else:
        tr.benchmark()
    if world_size > 1:
        dist.destroy_process_group()


def parse_gpus(s):
    """'0,1,2' or '0-7' (run_swin_mae3d.py:857-864)."""
    if not s:
        return [0]
    out = []
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-")
            out += list(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return out


def main(argv=None):
    args = parse_args(argv)
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:      # launched by torchrun
        rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        main_worker(rank, world, list(range(world)), args, int(os.environ.get("MASTER_PORT", "29500")))
        return
    gpus = parse_gpus(args.gpus)
    if len(gpus) <= 1:
        main_worker(0, 1, gpus, args, 0)
    else:
        port = random.randint(20000, 60000)
        mp.spawn(main_worker, nprocs=len(gpus), args=(len(gpus), gpus, args, port))


if __name__ == "__main__":
    main()
