"""The training step of the reference on the CUDA path.

Reference: `AudioLightningModule.training_step` (system/audio_litmodule.py:83-124: est = model(mix);
loss = PITLossWrapper(...)(est, targets)) and what the Lightning `Trainer` does around it
(audio_train.py:71,187-197): backward, DDP gradient mean, `gradient_clip_val=5.0` (global norm),
`torch.optim.Adam` (configs/tdanet_lsr2.yml:42-45).

Here one `TrainingStep.step(mixtures, targets)` enqueues, on the current stream and without a host
round trip: tdanet_forward_train -> tdanet_pit_loss (value + d loss/d est) -> tdanet_backward ->
one NCCL all-reduce of the flat fp32 gradient (only when world_size > 1) -> tdanet_grad_sqnorm ->
tdanet_adam_step.  Parameters, gradients and the Adam moments live in flat buffers; every
`nn.Parameter` of the model (and its `.grad`) is a view into them, so `state_dict()`, checkpoints and
`torch.optim` interoperate unchanged.

Dropout / DropPath (p = 0.1 in the reference's train mode) are applied while
`model.training`: the keep-masks are drawn on the device (Philox, csrc/dropout.cu) inside the same enqueue, so
a CUDA-graph replay of the step draws fresh masks.  `model.dropout = model.drop_path = 0.0` (or `.eval()`)
gives the deterministic step the parity tests compare with autograd of the oracle.
"""
from __future__ import annotations

import weakref
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from ... import _lib
from ...engine import adam_step, grad_sqnorm, pit_loss


class FlatParameters:
    """Re-homes the parameters of `model` into one flat fp32 buffer (state_dict order) with a gradient
    buffer of the same layout; `p.data` / `p.grad` become views."""

    def __init__(self, model: torch.nn.Module):
        params = [(n, p) for n, p in model.named_parameters()]
        if not params:
            raise ValueError("model has no parameters")
        dev = params[0][1].device
        if dev.type != "cuda":
            raise _lib.TdanetError("training runs on CUDA parameters only; call model.cuda() first (no CPU path)")
        self.names: List[str] = [n for n, _ in params]
        sizes = [p.numel() for _, p in params]
        # 16-byte aligned slices so that vector loads of any parameter stay aligned
        offs, total = [], 0
        for s in sizes:
            offs.append(total)
            total += (s + 3) // 4 * 4
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.views: Dict[str, torch.Tensor] = {}
        self.grad_views: Dict[str, torch.Tensor] = {}
        for (n, p), o, s in zip(params, offs, sizes):
            v = self.flat[o:o + s].view(p.shape)
            v.copy_(p.data)
            p.data = v
            g = self.grad[o:o + s].view(p.shape)
            p.grad = g
            self.views[n] = v
            self.grad_views[n] = g

    def zero_grad(self) -> None:
        self.grad.zero_()


class TrainingStep:
    """forward + PIT loss + backward + gradient all-reduce + clip + Adam for a TDANetBest on this rank's GPU.

    model            tdanet_b200.look2hear.models.TDANetBest on a CUDA device
    loss             tdanet_b200.look2hear.losses.PITLossWrapper (pw_mtx)
    lr, betas, eps   torch.optim.Adam arguments (weight_decay 0 as in the reference config)
    clip_grad_norm   Trainer(gradient_clip_val=...) - 5.0 in audio_train.py:193; 0 disables
    process_group    torch.distributed group for the gradient mean (None: default group if initialised)
    """

    def __init__(self, model, loss, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 clip_grad_norm: float = 5.0, process_group=None):
        self.model = model
        self.loss = loss
        self.lr, self.betas, self.eps, self.clip = lr, betas, eps, clip_grad_norm
        self.group = process_group
        self.params = FlatParameters(model)
        dev = self.params.flat.device
        self.exp_avg = torch.zeros_like(self.params.flat)
        self.exp_avg_sq = torch.zeros_like(self.params.flat)
        self.step_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.sqnorm = torch.zeros(2, dtype=torch.float64, device=dev)
        self._w = self._gw = None
        self._graph = None
        self._graph_shapes = None
        # weakly: the engine outlives this object (a second TrainingStep for the same model must not keep the first
        # one's flat buffers and graph alive through the listener list)
        drop = weakref.WeakMethod(self._drop_graph)
        model.engine.on_train_workspace_realloc(lambda device, _drop=drop: (_drop() or (lambda d: None))(device))
        self.broadcast_state()

    def _drop_graph(self, device=None) -> None:
        """The captured graph bakes the training-workspace pointer in; the engine calls this before it replaces the
        workspace (a later eager step with a larger batch), so a stale replay can never write to freed memory."""
        self._graph = None
        self._graph_shapes = None

    def broadcast_state(self, src: int = 0) -> None:
        """What DistributedDataParallel's constructor does in the reference run (Lightning DDP, audio_train.py:187-197):
        every rank starts from rank `src`'s parameters - here also its Adam moments and step counter, so a checkpoint
        loaded on one rank resumes identically everywhere.  No-op in a single process."""
        if self._world() <= 1:
            return
        for t in (self.params.flat, self.exp_avg, self.exp_avg_sq, self.step_count):
            dist.broadcast(t, src=dist.get_global_rank(self.group, src) if self.group is not None else src,
                           group=self.group)

    # ------------------------------------------------------------------ pieces
    def _world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def _pack(self):
        if self._w is None:
            eng = self.model.engine
            sd = {k: v.detach() for k, v in self.model.state_dict(keep_vars=True).items()}
            self._w = eng.pack(sd)
            self._gw = eng.pack(dict(self.params.grad_views), optional=True)
        return self._w, self._gw

    def forward_backward(self, mixtures: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        """Enqueues forward, loss and backward; gradients are ADDED to the flat gradient buffer.
        Returns the loss as a 1-element CUDA tensor (no sync)."""
        if mixtures.ndim == 3:
            mixtures = mixtures.squeeze(1)
        wav = mixtures.float().contiguous()
        eng = self.model.engine
        w, gw = self._pack()
        self.model._sync_dropout()
        est = eng.forward_train(w, wav, self.model.attn_group)
        loss, _, _, d_est = pit_loss(est, targets, self.loss.loss_func.sdr_type, self.loss.threshold_byloss,
                                     want_grad=True)
        eng.backward(w, gw, wav, d_est, self.model.attn_group)
        return loss

    def optimizer_step(self) -> None:
        """DDP gradient mean (sum all-reduce, 1/world folded into the update), clip, Adam."""
        world = self._world()
        if world > 1:
            dist.all_reduce(self.params.grad, op=dist.ReduceOp.SUM, group=self.group)
        sq = None
        if self.clip and self.clip > 0:
            grad_sqnorm(self.params.grad, self.sqnorm)
            sq = self.sqnorm
        adam_step(self.params.flat, self.params.grad, self.exp_avg, self.exp_avg_sq, self.step_count, self.lr,
                  self.betas, self.eps, float(self.clip or 0.0), 1.0 / world, sq)

    # ------------------------------------------------------------------ the step
    def step(self, mixtures: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        """One optimisation step; returns the (pre-update) loss as a CUDA tensor without synchronising."""
        self.params.zero_grad()
        loss = self.forward_backward(mixtures, targets)
        self.optimizer_step()
        return loss

    # ------------------------------------------------------------------ CUDA graph
    def capture(self, mixtures: torch.Tensor, targets: torch.Tensor) -> None:
        """Captures zero-grad + forward + loss + backward (about 2k launches) for this batch shape as one CUDA
        graph; later `step()` calls with the same shapes copy into its static inputs and replay it.  The gradient
        all-reduce and the optimiser kernels stay outside the graph (NCCL + 3 launches)."""
        if mixtures.ndim == 3:
            mixtures = mixtures.squeeze(1)
        dev = self.params.flat.device
        self._static_mix = mixtures.float().contiguous().clone()
        self._static_tgt = targets.float().contiguous().clone()
        s = torch.cuda.Stream(dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):   # warm-up outside capture: workspaces, function attributes
            self.params.zero_grad()
            self.forward_backward(self._static_mix, self._static_tgt)
        torch.cuda.current_stream(dev).wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.params.zero_grad()
            self._static_loss = self.forward_backward(self._static_mix, self._static_tgt)
        self._graph = g
        self._graph_shapes = self._graph_key(self._static_mix, self._static_tgt)

    def _graph_key(self, mixtures: torch.Tensor, targets: torch.Tensor):
        """Everything the captured step bakes in besides the buffers: shapes and the modes the library reads from
        the configuration at enqueue time."""
        m = self.model
        return (tuple(mixtures.shape), tuple(targets.shape), m.training, m.attn_group, m.gemm_mode, m.act_dtype,
                float(m.dropout), float(m.drop_path), self.loss.loss_func.sdr_type, bool(self.loss.threshold_byloss))

    def step_captured(self, mixtures: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        if mixtures.ndim == 3:
            mixtures = mixtures.squeeze(1)
        shapes = self._graph_key(mixtures, targets)
        if self._graph is None or self._graph_shapes != shapes:
            self.capture(mixtures, targets)       # first call, new batch shape / mode, or the workspace moved
        self._static_mix.copy_(mixtures, non_blocking=True)
        self._static_tgt.copy_(targets, non_blocking=True)
        self._graph.replay()
        self.optimizer_step()
        return self._static_loss

    # ------------------------------------------------------------------ checkpoints (reference Lightning layout)
    def _named_shapes(self):
        return [(n, self.params.views[n].shape) for n in self.params.names]

    def optimizer_state_dict(self):
        """`torch.optim.Adam(model.parameters()).state_dict()` of this run (audio_train.py:208-213 stores it under
        `optimizer_states`)."""
        from .checkpoint import adam_state_to_torch
        return adam_state_to_torch(self._named_shapes(), self.exp_avg, self.exp_avg_sq, int(self.step_count.item()),
                                   self.lr, self.betas, self.eps)

    def load_optimizer_state_dict(self, state) -> None:
        from .checkpoint import torch_to_adam_state
        step, self.lr, self.betas, self.eps = torch_to_adam_state(self._named_shapes(), state, self.exp_avg,
                                                                  self.exp_avg_sq)
        self.step_count.fill_(step)

    def checkpoint(self, epoch: int = 0, schedulers=()):
        """A dict laid out like the reference's Lightning checkpoint (loadable by `BaseModel.from_pretrain`)."""
        from .checkpoint import lightning_state_dict
        return {"epoch": epoch, "global_step": int(self.step_count.item()),
                "state_dict": lightning_state_dict({k: v.detach().clone() for k, v in self.model.state_dict().items()}),
                "optimizer_states": [self.optimizer_state_dict()],
                "lr_schedulers": [s.state_dict() for s in schedulers]}

    def load_checkpoint(self, ckpt, schedulers=()) -> None:
        from .checkpoint import strip_lightning_prefix
        self.model.load_state_dict(strip_lightning_prefix(ckpt["state_dict"]))   # copies into the flat views
        if ckpt.get("optimizer_states"):
            self.load_optimizer_state_dict(ckpt["optimizer_states"][0])
        for s, st in zip(schedulers, ckpt.get("lr_schedulers", [])):
            s.load_state_dict(st)
        self.broadcast_state()

    def grad_norm(self) -> float:
        """Global L2 norm of the gradient of the last step (after the all-reduce, before scaling); syncs."""
        return float(self.sqnorm[0].sqrt().item())


class AudioLightningModule(torch.nn.Module):
    """The part of the reference's LightningModule that is on the hot path (audio_litmodule.py:36-124):
    same constructor keywords, `forward(wav)`, `training_step(batch, batch_nb) -> {"loss": loss}`.
    `fit_step(batch)` additionally does what the Lightning Trainer does after `training_step`
    (backward, gradient mean, clip, Adam) through `TrainingStep`."""

    def __init__(self, audio_model=None, video_model=None, optimizer=None, loss_func=None, train_loader=None,
                 val_loader=None, test_loader=None, scheduler=None, config=None, log_freq=100):
        super().__init__()
        self.audio_model = audio_model
        self.video_model = video_model
        self.optimizer = optimizer
        self.loss_func = loss_func
        self.train_loader, self.val_loader, self.test_loader = train_loader, val_loader, test_loader
        self.scheduler = scheduler
        self.config = {} if config is None else config
        self.log_freq = log_freq
        self.default_monitor = "val_loss/dataloader_idx_0"
        self._step: Optional[TrainingStep] = None

    def forward(self, wav, mouth=None):
        return self.audio_model(wav)

    def training_step(self, batch, batch_nb=0):
        mixtures, targets = batch[0], batch[1]
        if self.config.get("training", {}).get("SpeedAug", False):
            raise NotImplementedError("SpeedAug (speechbrain SpeedPerturb) is a data augmentation outside the hot path")
        est_sources = self(mixtures)
        loss = self.loss_func["train"](est_sources, targets)
        return {"loss": loss}

    def validation_step(self, batch, batch_nb=0, dataloader_idx=0):
        """audio_litmodule.py:127-165: PIT loss of the separated sources under no_grad (eval forward)."""
        mixtures, targets = batch[0], batch[1]
        was_training = self.audio_model.training
        self.audio_model.eval()
        with torch.no_grad():
            loss = self.loss_func["val"](self(mixtures), targets)
        self.audio_model.train(was_training)
        return {"val_loss" if dataloader_idx == 0 else "test_loss": loss}

    def fit_step(self, batch, lr: Optional[float] = None, clip_grad_norm: float = 5.0):
        mixtures, targets = batch[0], batch[1]
        if self._step is None:
            opt = self.config.get("optimizer", {})
            self._step = TrainingStep(self.audio_model, self.loss_func["train"],
                                      lr=lr if lr is not None else opt.get("lr", 1e-3), clip_grad_norm=clip_grad_norm)
        return {"loss": self._step.step(mixtures, targets)}
