"""Training-loop semantics of the reference's seg trainers around the DCNv3 hot path,
without their logging / data / validation code (out of scope):

  hyper-parameters ........ unet-lite/yolo5-seg/seg_diceloss_yolov5.py:851-862 (lr0 0.01, lrf 0.2, momentum
                            0.937, weight_decay 5e-4)
  nominal batch 64 ........ :970-972   accumulate = max(round(64 / batch), 1); weight_decay *= batch*accumulate/64
  optimizer ............... utils/torch_utils.py:318-346 (smart_optimizer; restated in seg.py)
  LR schedule ............. :975-980   linear to lrf, or one_cycle cosine (utils/general.py one_cycle)
  EMA ..................... utils/torch_utils.py:404-432 (ModelEMA, decay 0.9999 with a tau=2000 ramp)
  hot loop ................ :1073-1103 forward (AMP) -> loss -> backward -> every `accumulate` micro-steps:
                            optimizer.step, zero_grad, ema.update
  checkpoint .............. :1204-1212 {'model': ema.ema, 'optimizer', 'epoch', 'best_fitness'}; best = {'model'}
  resume .................. utils/torch_utils.py:361-378 (smart_resume)

Differences, on purpose: autocast dtype is a parameter (bf16 needs no GradScaler; fp16 uses one, as the
reference's --amp does, :1043-1055); under DDP the non-boundary micro-steps run inside `no_sync()` so the
gradient all-reduce happens once per optimizer step (the reference never runs its seg trainers under DDP).
"""
from __future__ import annotations

import math
import warnings
from contextlib import nullcontext
from copy import deepcopy
from typing import Optional

import torch
from torch import nn

from .seg import forward_loss, smart_optimizer, wrap_ddp

HYP = dict(lr0=0.01, lrf=0.2, momentum=0.937, weight_decay=0.0005, label_smoothing=0.0)
NOMINAL_BATCH = 64


def de_parallel(model: nn.Module) -> nn.Module:
    return model.module if isinstance(model, (nn.parallel.DataParallel, nn.parallel.DistributedDataParallel)) else model


def lr_lambda(epochs: int, lrf: float = HYP["lrf"], cos: bool = False):
    """Epoch -> LR multiplier.  Linear: (1 - x/epochs)(1 - lrf) + lrf (seg_diceloss_yolov5.py:979);
    cosine: one_cycle(1, lrf, epochs) = ((1 - cos(x*pi/epochs))/2)(lrf - 1) + 1 (:977)."""
    if cos:
        return lambda x: ((1 - math.cos(x * math.pi / epochs)) / 2) * (lrf - 1) + 1
    return lambda x: (1 - x / epochs) * (1.0 - lrf) + lrf


class ModelEMA:
    """Exponential moving average of every floating-point entry of the model's state_dict, kept in the
    model's own precision, decay ramped as d(u) = decay * (1 - exp(-u / tau)) (torch_utils.py:404-426)."""

    def __init__(self, model: nn.Module, decay: float = 0.9999, tau: float = 2000, updates: int = 0):
        self.ema = deepcopy(de_parallel(model)).eval()
        self.updates = updates
        self.decay = lambda x: decay * (1 - math.exp(-x / tau))
        for p in self.ema.parameters():
            p.requires_grad_(False)

    @torch.no_grad()
    def update(self, model: nn.Module) -> None:
        self.updates += 1
        d = self.decay(self.updates)
        msd = de_parallel(model).state_dict()
        for k, v in self.ema.state_dict().items():
            if v.dtype.is_floating_point:
                v.mul_(d).add_(msd[k].detach(), alpha=1 - d)


class Trainer:
    """One object per process (one process per GPU).  `micro_step` is the body of the reference's inner loop."""

    def __init__(self, model: nn.Module, criterion: nn.Module, batch_size: int, epochs: int = 300,
                 hyp: Optional[dict] = None, cos_lr: bool = False, autocast_dtype=None, device=None,
                 ema: bool = True, sync_bn: bool = False):
        self.hyp = dict(HYP, **(hyp or {}))
        if sync_bn:  # --sync-bn, seg_diceloss_yolov5.py:991-992 (only meaningful with a process group)
            model = nn.SyncBatchNorm.convert_sync_batchnorm(model)
            if device is not None:
                model = model.to(device)
        self.raw_model = model
        self.accumulate = max(round(NOMINAL_BATCH / batch_size), 1)
        # found by running this class under NCCL (bench.py trainer_smoke): DDP(static_graph=True) asserts when its first
        # backward runs inside no_sync(), which is exactly what gradient accumulation does
        self.model = wrap_ddp(model, device, static_graph=self.accumulate == 1)
        self.criterion = criterion
        self.epochs = epochs
        wd = self.hyp["weight_decay"] * batch_size * self.accumulate / NOMINAL_BATCH
        self.optimizer = smart_optimizer(self.model, self.hyp["lr0"], self.hyp["momentum"], wd)
        self.scheduler = torch.optim.lr_scheduler.LambdaLR(self.optimizer, lr_lambda(epochs, self.hyp["lrf"], cos_lr))
        self.ema = ModelEMA(self.model) if ema else None
        self.autocast_dtype = autocast_dtype
        self.scaler = torch.amp.GradScaler("cuda") if autocast_dtype == torch.float16 else None
        self.epoch, self.best_fitness, self._i = 0, 0.0, 0
        self.optimizer.zero_grad(set_to_none=True)

    def micro_step(self, imgs: torch.Tensor, labels: torch.Tensor, last_of_epoch: bool = False):
        """forward -> loss -> backward; optimizer / EMA update every `accumulate` calls (or at the epoch's
        last batch, seg_diceloss_yolov5.py:1095).  Returns (loss, (total, ce, dice), stepped)."""
        boundary = (self._i + 1) % self.accumulate == 0 or last_of_epoch
        sync = nullcontext() if boundary or not hasattr(self.model, "no_sync") else self.model.no_sync()
        with sync:
            loss, parts = forward_loss(self.model, self.criterion, imgs, labels, self.autocast_dtype)
            (self.scaler.scale(loss) if self.scaler else loss).backward()
        if boundary:
            if self.scaler:
                self.scaler.step(self.optimizer)
                self.scaler.update()
            else:
                self.optimizer.step()
            self.optimizer.zero_grad(set_to_none=True)
            if self.ema:
                self.ema.update(self.model)
        self._i = 0 if last_of_epoch else self._i + 1
        return loss.detach(), parts, boundary

    def end_epoch(self, fitness: Optional[float] = None) -> None:
        self.scheduler.step()
        if fitness is not None and fitness > self.best_fitness:
            self.best_fitness = fitness
        self.epoch += 1

    # ---- checkpoints: the reference pickles the whole EMA module (seg_diceloss_yolov5.py:1204-1212)
    def checkpoint(self) -> dict:
        """Call after `end_epoch()` (the reference saves after its scheduler step, :1204-1212): 'epoch' is the index
        of the epoch just FINISHED, so `resume` continues at that index + 1 and the schedule is stepped once per
        epoch actually trained."""
        m = self.ema.ema if self.ema else de_parallel(self.model)
        return {"model": deepcopy(m), "optimizer": self.optimizer.state_dict(), "epoch": self.epoch - 1,
                "best_fitness": self.best_fitness}

    def save(self, last_path, best_path=None, is_best: bool = False) -> None:
        ck = self.checkpoint()
        torch.save(ck, last_path)
        if best_path is not None and is_best:
            torch.save({"model": ck["model"]}, best_path)

    def resume(self, ckpt: dict) -> None:
        """smart_resume (torch_utils.py:361-378) + the partial weight load of :944-952 (intersect by name/shape)."""
        src = ckpt["model"].float().state_dict()
        dst = de_parallel(self.model).state_dict()
        keep = {k: v for k, v in src.items() if k in dst and v.shape == dst[k].shape}
        bad = [k for k, v in src.items() if k in dst and v.shape != dst[k].shape]
        if bad:  # the reference's intersect drops these silently (:944-952); a dropped DCNv3 head stays zero-initialised
            warnings.warn(f"Trainer.resume: {len(bad)} checkpoint tensors do not fit this model's shapes and were NOT "
                          f"loaded (first: {bad[0]} {tuple(src[bad[0]].shape)} vs {tuple(dst[bad[0]].shape)}); build the "
                          f"blocks with the checkpoint's DCNv3 group count (dcn_group=...)", RuntimeWarning, stacklevel=2)
        self.dropped_keys = bad
        de_parallel(self.model).load_state_dict(keep, strict=False)
        if ckpt.get("optimizer") is not None:
            self.optimizer.load_state_dict(ckpt["optimizer"])
            self.best_fitness = ckpt.get("best_fitness", 0.0)
        self.epoch = ckpt.get("epoch", -1) + 1
        for _ in range(self.epoch):
            self.scheduler.step()
        if self.ema:
            self.ema.ema.load_state_dict(src, strict=False)
