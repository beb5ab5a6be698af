"""The slice of pytorch_lightning.Trainer the reference actually uses (train.py:475-509): epochs over the bucketed
loader, gradient accumulation (inside W2V2Distil.training_step), the s3prl warm-up / linear-decay schedule sized from
the loader (train.py:411-413), a validation pass per epoch, ModelCheckpoint(save_last, save_top_k=3, monitor v_loss),
EarlyStopping(patience 15) and resume.  Checkpoints keep Lightning's shape ({'state_dict': {'student_model.*'}, ...})
so the reference's own UpstreamExpert (fithubert/expert.py:40-43) loads them."""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from .checkpoint import load_checkpoint_to_cpu, load_student_state_dict, student_checkpoint


def total_training_steps(n_batches: int, num_gpus: int, num_epochs: int, accumulate: int) -> int:
    """train.py:411-412: (num_epochs * (len(loader) // gpus)) // accumulate_grad_batches."""
    return (num_epochs * (n_batches // max(1, num_gpus))) // max(1, accumulate)


class TopKCheckpoints:
    """ModelCheckpoint(dirpath, filename='checkpoint-{epoch:02d}', save_last=True, save_top_k=3, monitor='v_loss',
    mode='min') (train.py:475-483)."""

    def __init__(self, dirpath: str, k: int = 3):
        self.dirpath, self.k = dirpath, k
        self.best: List[tuple] = []  # (v_loss, path)

    def save(self, state: Dict, epoch: int, v_loss: float) -> Optional[str]:
        os.makedirs(self.dirpath, exist_ok=True)
        torch.save(state, os.path.join(self.dirpath, "last.ckpt"))
        path = os.path.join(self.dirpath, f"checkpoint-epoch={epoch:02d}.ckpt")
        if len(self.best) < self.k or v_loss < max(b[0] for b in self.best):
            torch.save(state, path)
            self.best.append((v_loss, path))
            self.best.sort(key=lambda b: b[0])
            for _, old in self.best[self.k:]:
                if os.path.exists(old):
                    os.remove(old)
            self.best = self.best[:self.k]
            return path
        return None


class EarlyStopping:
    """EarlyStopping(monitor='v_loss', patience=15, mode='min') (train.py:485-490)."""

    def __init__(self, patience: int = 15, min_delta: float = 0.0):
        self.patience, self.min_delta, self.best, self.bad = patience, min_delta, float("inf"), 0

    def step(self, value: float) -> bool:
        if value < self.best - self.min_delta:
            self.best, self.bad = value, 0
        else:
            self.bad += 1
        return self.bad >= self.patience


def checkpoint_state(model, epoch: int, global_step: int, v_loss: float) -> Dict:
    opt = model.optimizer
    state = student_checkpoint(model.student_model, {"epoch": epoch, "global_step": global_step, "v_loss": v_loss})
    if opt is not None and getattr(opt, "m", None) is not None:
        state["optimizer_states"] = [{"m": opt.m.cpu(), "v": opt.v.cpu(), "step_count": opt.step_count,
                                      "total_steps": opt.total_steps}]
    return state


def restore(model, path: str) -> Dict:
    state = load_checkpoint_to_cpu(path)
    model.student_model.load_state_dict(load_student_state_dict(state))
    if model.optimizer is not None and state.get("optimizer_states"):
        o = state["optimizer_states"][0]
        model.optimizer._build()
        model.optimizer.m.copy_(o["m"])
        model.optimizer.v.copy_(o["v"])
        model.optimizer.step_count = int(o["step_count"])
    return state


def _global_mean(total, count: int, world: int) -> float:
    """sum(total) / sum(count) over all ranks (one small all-reduce; NaN when no rank saw a batch)."""
    if world > 1:
        dev = total.device if isinstance(total, torch.Tensor) else (
            torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
        pair = torch.zeros(2, dtype=torch.float64, device=dev)
        if total is not None:
            pair[0] = float(total) if not isinstance(total, torch.Tensor) else total.detach().double()
        pair[1] = float(count)
        dist.all_reduce(pair)
        total, count = pair[0], float(pair[1])
    if total is None or count <= 0:
        return float("nan")
    return float(total) / float(count)


def fit(model, train_loader, val_loader=None, num_epochs: Optional[int] = None, output_dir: Optional[str] = None,
        ckpt_path: Optional[str] = None, save_top_k: int = 3, patience: int = 15, log_every: int = 0) -> Dict:
    """trainer.fit(model, ckpt_path=...) of train.py:492-509.  `model` is a W2V2Distil; the loaders are BucketLoaders
    (one per rank).  Returns {'epochs_run', 'global_step', 'history': [(epoch, train_loss, v_loss)], 'stopped_early'}."""
    t = model.train_cfg
    num_epochs = int(num_epochs if num_epochs is not None else t["num_epochs"])
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    n_batches = len(train_loader.dataset) if hasattr(train_loader, "dataset") else len(train_loader)
    model.configure_optimizers(total_steps=total_training_steps(n_batches, world, num_epochs, model.accumulate))
    start_epoch, global_step = 0, 0
    if ckpt_path:
        st = restore(model, ckpt_path)
        start_epoch, global_step = int(st.get("epoch", -1)) + 1, int(st.get("global_step", 0))
    saver = TopKCheckpoints(output_dir, save_top_k) if (output_dir and rank == 0) else None
    stopper = EarlyStopping(patience)
    history, stopped = [], False
    for epoch in range(start_epoch, num_epochs):
        if hasattr(train_loader, "set_epoch"):
            train_loader.set_epoch(epoch)
        model.student_model.train()
        run, nb = None, 0
        for i, batch in enumerate(train_loader):
            loss = model.training_step(batch, i)
            run = loss.detach() if run is None else run + loss.detach()  # stays on the device: no per-step sync
            nb += 1
            if model._micro == 0:
                global_step += 1
            if log_every and rank == 0 and (i + 1) % log_every == 0:
                print(f"epoch {epoch} step {i + 1}: loss {float(loss):.5f} lr {model.optimizer.current_lr():.3e}", flush=True)
        model.training_epoch_end()
        # every rank sees only its shard: the monitored values are (sum, count) pairs reduced over the ranks, so all
        # ranks rank checkpoints and hit EarlyStopping on the same epoch (Lightning syncs both; a rank-local decision
        # would leave the others blocked in the next gradient all-reduce)
        train_loss = _global_mean(run, nb, world)
        v_loss = float("nan")
        if val_loader is not None:
            model.student_model.eval()
            tot, cnt = None, 0
            for i, batch in enumerate(val_loader):
                v = model.validation_step(batch, i)["v_loss"] * batch["x"].shape[0]
                tot = v if tot is None else tot + v
                cnt += batch["x"].shape[0]
            v_loss = _global_mean(tot, cnt, world)
        history.append((epoch, train_loss, v_loss))
        monitor = v_loss if val_loader is not None else train_loss
        if saver is not None:
            saver.save(checkpoint_state(model, epoch, global_step, monitor), epoch, monitor)
        if stopper.step(monitor):
            stopped = True
            break
    return {"epochs_run": len(history), "global_step": global_step, "history": history, "stopped_early": stopped}
