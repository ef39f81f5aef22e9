"""Checkpoint interchange with the reference's Lightning checkpoints and its plateau scheduler.

Reference: audio_train.py:92-99 (ReduceLROnPlateau(factor=0.5, patience=...)), :145-154 / :208-213 (ModelCheckpoint
writes `{epoch, global_step, state_dict: {"audio_model.<key>": ...}, optimizer_states: [Adam.state_dict()],
lr_schedulers: [...]}`; best_model.pth = `serialize()`), models/base_model.py:135-173 (`from_pretrain` strips the
`audio_model.` prefix).  The fused training step keeps Adam's moments in flat buffers; the functions here convert
between that layout and `torch.optim.Adam.state_dict()` so that either side can resume the other's run.
Pure host logic on tensors of any device.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch

PREFIX = "audio_model."


def param_layout(named_shapes: Sequence[Tuple[str, torch.Size]]):
    """Offsets of every parameter in the flat buffers of `FlatParameters` (16-byte aligned slices)."""
    offs, total = [], 0
    for _, shape in named_shapes:
        n = int(torch.Size(shape).numel())
        offs.append(total)
        total += (n + 3) // 4 * 4
    return offs, total


def adam_state_to_torch(named_shapes, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: int, lr: float,
                        betas=(0.9, 0.999), eps: float = 1e-8) -> Dict:
    """Flat moments -> `torch.optim.Adam(model.parameters()).state_dict()` (parameters in named_parameters order)."""
    offs, _ = param_layout(named_shapes)
    state = {}
    for i, ((_, shape), o) in enumerate(zip(named_shapes, offs)):
        n = int(torch.Size(shape).numel())
        if step > 0:
            state[i] = {"step": torch.tensor(float(step)), "exp_avg": exp_avg[o:o + n].reshape(shape).clone(),
                        "exp_avg_sq": exp_avg_sq[o:o + n].reshape(shape).clone()}
    group = {"lr": lr, "betas": tuple(betas), "eps": eps, "weight_decay": 0, "amsgrad": False, "maximize": False,
             "foreach": None, "capturable": False, "differentiable": False, "fused": None,
             "params": list(range(len(named_shapes)))}
    return {"state": state, "param_groups": [group]}


def torch_to_adam_state(named_shapes, opt_state: Dict, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor):
    """`Adam.state_dict()` -> flat moments (in place); returns (step, lr, betas, eps).  Parameters without state
    (never updated, e.g. the dead loc_glo_fus of the last scale) get zero moments."""
    offs, _ = param_layout(named_shapes)
    exp_avg.zero_()
    exp_avg_sq.zero_()
    step = 0
    for i, ((_, shape), o) in enumerate(zip(named_shapes, offs)):
        st = opt_state["state"].get(i)
        if st is None:
            continue
        n = int(torch.Size(shape).numel())
        exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
        exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
        step = max(step, int(st["step"]))
    g = opt_state["param_groups"][0]
    return step, g["lr"], tuple(g["betas"]), g["eps"]


def lightning_state_dict(model_state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """state_dict of the model -> the `state_dict` entry of a Lightning checkpoint of AudioLightningModule."""
    return {PREFIX + k: v for k, v in model_state.items()}


def strip_lightning_prefix(state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """The inverse, as `BaseModel.from_pretrain` does for local files (base_model.py:137-148)."""
    return {(k[len(PREFIX):] if k.startswith(PREFIX) else k): v for k, v in state.items()}


class ReduceLROnPlateau:
    """torch.optim.lr_scheduler.ReduceLROnPlateau semantics (mode "min", relative threshold) for an object with an
    `lr` attribute (TrainingStep); state_dict keys match torch's so the `lr_schedulers` entry interchanges."""

    def __init__(self, optimizer, factor: float = 0.5, patience: int = 10, threshold: float = 1e-4, cooldown: int = 0,
                 min_lr: float = 0.0, eps: float = 1e-8):
        if factor >= 1.0:
            raise ValueError("Factor should be < 1.0.")
        self.optimizer = optimizer
        self.factor, self.patience, self.threshold, self.cooldown, self.min_lr, self.eps = \
            factor, patience, threshold, cooldown, min_lr, eps
        self.mode, self.threshold_mode = "min", "rel"
        self.best = float("inf")
        self.num_bad_epochs = 0
        self.cooldown_counter = 0
        self.last_epoch = 0

    def step(self, metric: float) -> None:
        current = float(metric)
        self.last_epoch += 1
        if current < self.best * (1.0 - self.threshold):
            self.best = current
            self.num_bad_epochs = 0
        else:
            self.num_bad_epochs += 1
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.num_bad_epochs = 0
        if self.num_bad_epochs > self.patience:
            new_lr = max(self.optimizer.lr * self.factor, self.min_lr)
            if self.optimizer.lr - new_lr > self.eps:
                self.optimizer.lr = new_lr
            self.cooldown_counter = self.cooldown
            self.num_bad_epochs = 0

    def state_dict(self) -> Dict:
        return {k: v for k, v in self.__dict__.items() if k != "optimizer"}

    def load_state_dict(self, state: Dict) -> None:
        self.__dict__.update({k: v for k, v in state.items() if k != "optimizer"})
