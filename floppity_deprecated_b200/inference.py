"""``SNPE_C`` / ``DirectPosterior`` / ``BoxUniform`` -- drop-ins for the sbi 0.22 objects FlopPITy drives:

    modules.py:57      utils.BoxUniform(low, high, device)
    modules.py:64      SNPE_C(prior, density_estimator, device)
    floppity.py:216    inference.append_simulations(theta, x, proposal)
    floppity.py:219    inference.train(show_train_summary, stop_after_epochs, num_atoms, force_first_round_loss,
                                       retrain_from_scratch, use_combined_loss)
    floppity.py:225    inference.build_posterior(net).set_default_x(x)
    floppity.py:252    proposal.sample((n,))            modules.py:102, floppityFUN.py:320
    floppityFUN.py:35  posterior.log_prob(theta, x=default_x)

Same names, arguments, defaults and error behaviour; the arithmetic underneath is the sm_100a kernel library.
The training step is a fixed launch sequence on preallocated device buffers (gather -> forward -> loss ->
hand-written backward -> [NCCL all-reduce] -> clip+Adam) with no host synchronisation inside an epoch.
"""
from __future__ import annotations

import copy
import time
import warnings
from typing import Callable, Optional

import numpy as np
import torch
from torch import Tensor

from . import _lib, dist as fdist
from .flow import NSFFlow, _valid_rows


# ---------------------------------------------------------------------------------------------------------
class BoxUniform:
    """``sbi.utils.BoxUniform(low, high, device)`` = Independent(Uniform(low, high), 1)  (modules.py:57)."""

    def __init__(self, low, high, reinterpreted_batch_ndims: int = 1, device: str = "cpu"):
        self.device = torch.device(device)
        self.low = torch.as_tensor(low, dtype=torch.float32).reshape(-1).to(self.device)
        self.high = torch.as_tensor(high, dtype=torch.float32).reshape(-1).to(self.device)
        if self.low.shape != self.high.shape:
            raise ValueError("low and high must have the same shape")

    @property
    def event_shape(self):
        return torch.Size([self.low.numel()])

    @property
    def mean(self):
        return (self.low + self.high) / 2

    def to(self, device):
        self.device = torch.device(device)
        self.low, self.high = self.low.to(device), self.high.to(device)
        return self

    def within(self, theta: Tensor) -> Tensor:
        """``prior.support.check``: all dims in [low, high]."""
        theta = torch.as_tensor(theta, dtype=torch.float32).to(self.low.device)
        return ((theta >= self.low) & (theta <= self.high)).all(dim=-1)

    def log_prob(self, theta) -> Tensor:
        theta = torch.as_tensor(theta, dtype=torch.float32).to(self.low.device)
        lp = -torch.log(self.high - self.low).sum()
        inside = ((theta >= self.low) & (theta < self.high)).all(dim=-1)
        return torch.where(inside, lp.expand(inside.shape), torch.full(inside.shape, -float("inf"), device=lp.device))

    def sample(self, sample_shape=torch.Size()) -> Tensor:
        shape = tuple(torch.Size(sample_shape)) + (self.low.numel(),)
        return self.low + torch.rand(shape, device=self.low.device) * (self.high - self.low)


def _check_fp16_range(st) -> None:
    if len(st) > 2 and st[2]:
        raise FloatingPointError(
            f"{st[2]} operand pieces of the conditioner kernels reached the edge of fp16's range (|value| > 6e4): the 3xFP16 "
            "operand format does not cover this flow's activations / gradients -- set FLOPPITY_FUSED_FP16=0 and "
            "FLOPPITY_GEMM_FP16=0 (3xTF32 operands) and check that the observations are z-scored")


# ---------------------------------------------------------------------------------------------------------
class _TrainEngine:
    """One optimiser step of ``PosteriorEstimator.train`` as a fixed kernel sequence on preallocated buffers."""

    def __init__(self, net: NSFFlow, prior: BoxUniform, batch: int, atomic: bool, num_atoms: int, combined: bool,
                 lr: float, clip_max_norm: Optional[float], use_graph: bool = False):
        # an embedding net in front of the flow (modules.py:21-45, -embedding FC) is trained jointly: its forward runs
        # before the flow's, its backward after, and both parameter sets share one clip norm and one Adam step count
        self.emb = getattr(net, "embedding", None)
        self.emb_trainable = self.emb is not None and self.emb.param_count > 0
        self.x_dim = net.C                           # width of the (standardised) observation rows
        if self.emb is not None:
            net = net.flow
        self.net, self.B = net, int(batch)
        dev = net._flat.device
        self.dev = dev
        self.atomic = bool(atomic)
        self.A = int(min(max(num_atoms, 2), batch)) if atomic else 1
        self.combined = bool(combined)
        self.lr, self.clip = float(lr), (float(clip_max_norm) if clip_max_norm is not None else 0.0)
        self.R = self.B * self.A
        f32 = dict(dtype=torch.float32, device=dev)
        self.grads = torch.zeros(net.grad_count, **f32)
        self.m = torch.zeros(net.param_count, **f32)
        self.v = torch.zeros(net.param_count, **f32)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.sumsq = torch.zeros(1, **f32)
        self.loss_sum = torch.zeros(1, **f32)
        self.prepared = torch.empty(net.prepared_count, **f32)
        self.ws = torch.empty(net.workspace_floats(self.R, self.B, True), **f32)
        self.logp = torch.empty(self.R, **f32)
        self.d_logp = torch.full((self.R,), -1.0 / self.B, **f32)
        self.theta_b = torch.empty(self.B, net.D, **f32)
        self.x_b = torch.empty(self.B, self.x_dim, **f32)
        if self.emb is not None:
            self.emb_grads = torch.zeros(max(self.emb.param_count, 4), **f32)
            self.emb_m = torch.zeros(self.emb.param_count, **f32)
            self.emb_v = torch.zeros(self.emb.param_count, **f32)
            self.d_ctx = torch.empty(self.B, net.C, **f32)
            self._emb_saved = None
        self.mask_b = torch.zeros(self.B, **f32)
        self.idx = torch.zeros(self.B, dtype=torch.int64, device=dev)
        self.status = torch.zeros(4, dtype=torch.int32, device=dev)
        if self.atomic:
            self.choices = torch.empty(self.B, self.A - 1, dtype=torch.int32, device=dev)
            self.atom_rows = torch.empty(self.R, net.D, **f32)
            self.out_lp = torch.empty(self.B, **f32)
            self.low = prior.low.to(dev).contiguous()
            self.high = prior.high.to(dev).contiguous()
        self.cfg = _lib.NsfConfig(**net.cfg)
        self.zlogdet = net._zlogdet()
        self.seed = 0
        self.world = fdist.rank_world()[1]
        self.graph = None
        self.use_graph = bool(use_graph) and self.world == 1 and net.cfg["dropout_p"] == 0.0 and self.emb is None
        self.launches_per_step = 0

    # -- pieces -------------------------------------------------------------------------------------------
    def _gather(self, theta_all, x_all, masks_all):
        L, s = _lib.lib(), _lib.stream()
        _lib.check(L.fp_gather_rows(_lib.ptr(theta_all), _lib.ptr(self.idx), _lib.ptr(self.theta_b), self.B, self.net.D, s), "gather theta")
        _lib.check(L.fp_gather_rows(_lib.ptr(x_all), _lib.ptr(self.idx), _lib.ptr(self.x_b), self.B, self.x_dim, s), "gather x")
        if masks_all is not None:
            _lib.check(L.fp_gather_rows(_lib.ptr(masks_all), _lib.ptr(self.idx), _lib.ptr(self.mask_b), self.B, 1, s), "gather masks")

    def _forward(self, training: bool):
        L, s, net, c = _lib.lib(), _lib.stream(), self.net, self.cfg
        drop_on = training and net.cfg["dropout_p"] > 0
        c.dropout_p = net.cfg["dropout_p"] if drop_on else 0.0
        if drop_on:
            self.seed += 1
        _lib.check(L.fp_nsf_prepare(c, _lib.ptr(net._flat), _lib.ptr(self.prepared), s), "prepare")
        if self.atomic:
            if self.use_graph:   # draw position from the device step counter: fresh atoms on every graph replay
                _lib.check(L.fp_atom_choices_step(_lib.ptr(self.choices), self.B, self.A, 12345, _lib.ptr(self.step_dev), s), "choices")
            else:
                _lib.check(L.fp_atom_choices(_lib.ptr(self.choices), self.B, self.A, self._choice_seed(), s), "choices")
            _lib.check(L.fp_atom_gather(_lib.ptr(self.theta_b), _lib.ptr(self.choices), _lib.ptr(self.atom_rows), self.B, self.A, net.D, s), "atom gather")
            theta = self.atom_rows
        else:
            theta = self.theta_b
        ctx = self.x_b
        if self.emb is not None:
            ctx, zs = self.emb.forward_raw(self.x_b, keep=training)
            self._emb_saved = (ctx, zs)
        self._ctx = ctx
        _lib.check(L.fp_nsf_logprob_forward(
            c, _lib.ptr(net._flat), _lib.ptr(self.prepared), _lib.ptr(theta), _lib.ptr(ctx), self.R, self.B,
            _lib.ptr(net._theta_shift), _lib.ptr(net._theta_scale), self.zlogdet, _lib.ptr(self.logp), _lib.ptr(self.ws),
            1, self.seed if drop_on else 0, _lib.ptr(self.status), s), "logprob forward")
        if self.atomic:
            _lib.check(L.fp_atomic_loss(_lib.ptr(self.logp), _lib.ptr(self.atom_rows), _lib.ptr(self.low), _lib.ptr(self.high),
                                        _lib.ptr(self.mask_b), int(self.combined), _lib.ptr(self.out_lp),
                                        _lib.ptr(self.d_logp) if training else None, self.B, self.A, net.D,
                                        _lib.ptr(self.status), s), "atomic loss")
            _lib.check(L.fp_sum(_lib.ptr(self.out_lp), self.B, -1.0, _lib.ptr(self.loss_sum), s), "loss sum")
        else:
            _lib.check(L.fp_sum(_lib.ptr(self.logp), self.B, -1.0, _lib.ptr(self.loss_sum), s), "loss sum")

    _choice_counter = 0

    def _choice_seed(self) -> int:
        # host-side stream position (plain launches); graph replays use the device step counter instead
        self._choice_counter += 1
        return (self._choice_counter * 0x9E3779B1 + 12345) & 0xFFFFFFFFFFFF

    def _backward_update(self):
        L, s, net, c = _lib.lib(), _lib.stream(), self.net, self.cfg
        n = net.param_count
        self.grads.zero_()
        self.sumsq.zero_()
        _lib.check(L.fp_nsf_logprob_backward(
            c, _lib.ptr(net._flat), _lib.ptr(self.prepared), _lib.ptr(self._ctx), self.R, self.B,
            _lib.ptr(net._theta_scale), _lib.ptr(self.d_logp), _lib.ptr(self.grads), None, _lib.ptr(self.ws),
            self.seed if c.dropout_p > 0 else 0, _lib.ptr(self.status), s), "logprob backward")
        _lib.check(L.fp_nsf_finish_grads(c, _lib.ptr(net._flat), _lib.ptr(self.prepared), _lib.ptr(self.grads), s), "finish grads")
        if self.emb_trainable:
            self.emb_grads.zero_()
            _lib.check(L.fp_nsf_ctx_backward(c, _lib.ptr(net._flat), self.R, self.B, _lib.ptr(self.ws), _lib.ptr(self.d_ctx), s),
                       "ctx backward")
            self.emb.backward_raw(self.x_b, self._emb_saved[1], self.d_ctx, self.emb_grads)
        scale = 1.0
        if self.world > 1:
            torch.cuda.nvtx.range_push("fp.allreduce")
            fdist.allreduce_sum_(self.grads[:n])
            torch.cuda.nvtx.range_pop()
            if self.emb_trainable:
                fdist.allreduce_sum_(self.emb_grads)
            scale = 1.0 / self.world
        # clip_grad_norm_ over ALL parameters (flow + embedding): both buffers feed one sum of squares
        _lib.check(L.fp_grad_sumsq(_lib.ptr(self.grads), n, _lib.ptr(self.sumsq), s), "sumsq")
        if self.emb_trainable:
            _lib.check(L.fp_grad_sumsq(_lib.ptr(self.emb_grads), self.emb.param_count, _lib.ptr(self.sumsq), s), "sumsq emb")
        self.step_dev += 1
        _lib.check(L.fp_adam_step(_lib.ptr(net._flat), _lib.ptr(self.grads), _lib.ptr(self.m), _lib.ptr(self.v), n,
                                  self.lr, 0.9, 0.999, 1e-8, _lib.ptr(self.step_dev), self.clip, scale,
                                  _lib.ptr(self.sumsq), s), "adam")
        if self.emb_trainable:
            _lib.check(L.fp_adam_step(_lib.ptr(self.emb._flat), _lib.ptr(self.emb_grads), _lib.ptr(self.emb_m),
                                      _lib.ptr(self.emb_v), self.emb.param_count, self.lr, 0.9, 0.999, 1e-8,
                                      _lib.ptr(self.step_dev), self.clip, scale, _lib.ptr(self.sumsq), s), "adam emb")

    # -- public -------------------------------------------------------------------------------------------
    def _step_body(self, theta_all, x_all, masks_all):
        nvtx = torch.cuda.nvtx
        nvtx.range_push("fp.gather")
        self._gather(theta_all, x_all, masks_all)
        nvtx.range_pop()
        nvtx.range_push("fp.forward+loss")
        self._forward(training=True)
        nvtx.range_pop()
        nvtx.range_push("fp.backward+allreduce+clip+adam")
        self._backward_update()
        nvtx.range_pop()

    def step(self, idx: Tensor, theta_all: Tensor, x_all: Tensor, masks_all: Optional[Tensor]):
        """One optimiser step on the rows ``idx`` of the resident dataset."""
        self.idx.copy_(idx)
        if self.use_graph:
            if self.graph is None:
                # warm-up on a side stream, then capture
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self._step_body(theta_all, x_all, masks_all)
                torch.cuda.current_stream().wait_stream(s)
                self._graph_data = (theta_all, x_all, masks_all)
                g = torch.cuda.CUDAGraph()
                n0 = _lib.lib().fp_launch_count()
                with torch.cuda.graph(g):
                    self._step_body(theta_all, x_all, masks_all)
                self.launches_per_step = int(_lib.lib().fp_launch_count() - n0)
                self.graph = g
            if self._graph_data[0] is not theta_all or self._graph_data[1] is not x_all:
                raise RuntimeError("graph was captured for a different resident dataset")
            self.graph.replay()
        else:
            self._step_body(theta_all, x_all, masks_all)

    def step_on_batch(self, theta_b: Tensor, x_b_std: Tensor, mask_b: Optional[Tensor] = None):
        """One optimiser step on an explicit batch already on the device (x standardised)."""
        self.theta_b.copy_(theta_b)
        self.x_b.copy_(x_b_std)
        if mask_b is not None:
            self.mask_b.copy_(mask_b.reshape(-1))
        self._forward(training=True)
        self._backward_update()

    @torch.no_grad()
    def eval_batch(self, idx: Tensor, theta_all, x_all, masks_all):
        """Accumulates the (negative) loss of a validation batch into ``loss_sum`` (no parameter update)."""
        self.idx.copy_(idx)
        self._gather(theta_all, x_all, masks_all)
        self._forward(training=False)

    def check_status(self):
        if self.world > 1:       # every rank must take the same decision, or the others hang in the next all-reduce
            stf = self.status.float()
            fdist.allreduce_sum_(stf)
            st = [int(v) for v in stf.tolist()]
        else:
            st = self.status.tolist()
        _check_fp16_range(st)
        if st[0] or st[1]:
            raise AssertionError(f"non-finite log-probs or atoms outside the prior: {st[0]}, negative spline discriminants: {st[1]}")


# ---------------------------------------------------------------------------------------------------------
class SNPE_C:
    """``sbi.inference.SNPE_C`` (APT) for the NSF estimator, as FlopPITy uses it (modules.py:64, floppity.py:216-225)."""

    def __init__(self, prior: Optional[BoxUniform] = None, density_estimator="nsf", device: str = "cpu",
                 logging_level="WARNING", summary_writer=None, show_progress_bars: bool = True):
        from .flow import posterior_nn
        self._prior = prior
        self._device = str(device)
        if isinstance(density_estimator, str):
            density_estimator = posterior_nn(model=density_estimator)
        self._build_neural_net: Callable = density_estimator
        self._neural_net: Optional[NSFFlow] = None
        self._theta_roundwise, self._x_roundwise, self._prior_masks = [], [], []
        self._data_round_index, self._proposal_roundwise = [], []
        self._round = 0
        self._val_log_prob = float("-inf")
        self._summary = dict(epochs_trained=[], best_validation_log_prob=[], validation_log_probs=[],
                             training_log_probs=[], epoch_durations_sec=[])
        self._engine = None

    # pickling (floppity.py:247 dumps the whole object): drop live engine buffers / graphs
    def __getstate__(self):
        d = self.__dict__.copy()
        d["_engine"] = None
        d.pop("_pinned", None)
        return d

    def append_simulations(self, theta, x, proposal=None, exclude_invalid_x: Optional[bool] = None,
                           data_device: Optional[str] = None):
        theta = torch.as_tensor(np.asarray(theta) if not torch.is_tensor(theta) else theta).detach().to("cpu", torch.float32)
        x = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).detach().to("cpu", torch.float32)
        if theta.shape[0] != x.shape[0]:
            raise ValueError("theta and x must have the same number of rows")
        if proposal is None or proposal is self._prior:
            current_round = 0
        else:
            current_round = (max(self._data_round_index) + 1) if self._data_round_index else 1
        if current_round == 0:
            ok = _valid_rows(x)
            if (~ok).any():
                warnings.warn(f"Found {int((~ok).sum())} NaN/Inf simulations; they are excluded from training.")
                theta, x = theta[ok], x[ok]
        elif not _valid_rows(x).all():
            raise ValueError("multi-round SNPE-C cannot drop NaN/Inf simulations; remove them before appending")
        masks = torch.ones(theta.shape[0], 1) if current_round == 0 else torch.zeros(theta.shape[0], 1)
        self._theta_roundwise.append(theta)
        self._x_roundwise.append(x)
        self._prior_masks.append(masks)
        self._proposal_roundwise.append(proposal)
        self._data_round_index.append(current_round)
        return self

    def _stage_rows(self, src: Tensor, rows: Tensor, key: str) -> Tensor:
        """``src[rows]`` gathered (multi-threaded) into a pinned buffer that is kept between ``train()`` calls, so the
        upload is one asynchronous DMA instead of a pageable staging copy."""
        cache = self.__dict__.setdefault("_pinned", {})
        shape = (rows.numel(),) + tuple(src.shape[1:])
        buf = cache.get(key)
        if buf is None or tuple(buf.shape) != shape:
            try:
                buf = torch.empty(shape, dtype=src.dtype, pin_memory=True)
            except RuntimeError:
                buf = torch.empty(shape, dtype=src.dtype)
            cache[key] = buf
        torch.index_select(src, 0, rows, out=buf)
        return buf

    def get_simulations(self, starting_round: int = 0):
        idx = [i for i, r in enumerate(self._data_round_index) if r >= starting_round]
        if len(idx) == 1:        # one round: hand the stored tensors over as they are (no host copy of a multi-GB dataset)
            i = idx[0]
            return self._theta_roundwise[i], self._x_roundwise[i], self._prior_masks[i]
        return (torch.cat([self._theta_roundwise[i] for i in idx]), torch.cat([self._x_roundwise[i] for i in idx]),
                torch.cat([self._prior_masks[i] for i in idx]))

    def _converged(self, epoch: int, stop_after_epochs: int) -> bool:
        converged = False
        net = self._neural_net
        if epoch == 0 or self._val_log_prob > self._best_val_log_prob:
            self._best_val_log_prob = self._val_log_prob
            self._epochs_since_last_improvement = 0
            # == deepcopy(state_dict) upstream (buffers never change in train()): EVERY trainable buffer, i.e. the flow's
            # flat parameters and, with an embedding net in front (-embedding FC, modules.py:21-45), the embedding's too
            self._best_state = [p.detach().clone() for p in net.parameters()]
        else:
            self._epochs_since_last_improvement += 1
        if self._epochs_since_last_improvement > stop_after_epochs - 1:
            with torch.no_grad():
                for p, best in zip(net.parameters(), self._best_state):
                    p.copy_(best)
            converged = True
        return converged

    def train(self, num_atoms: int = 10, training_batch_size: int = 50, learning_rate: float = 5e-4,
              validation_fraction: float = 0.1, stop_after_epochs: int = 20, max_num_epochs: int = 2 ** 31 - 1,
              clip_max_norm: Optional[float] = 5.0, calibration_kernel=None, resume_training: bool = False,
              force_first_round_loss: bool = False, discard_prior_samples: bool = False,
              use_combined_loss: bool = False, retrain_from_scratch: bool = False, show_train_summary: bool = False,
              dataloader_kwargs=None, seed: Optional[int] = None, use_cuda_graph: bool = True) -> NSFFlow:
        if calibration_kernel is not None:
            raise NotImplementedError("calibration kernels are not on the FlopPITy path (kernel == 1)")
        dev = _lib.require_cuda(None if self._device in ("cpu", "cuda") else self._device)
        if self._device == "cpu":
            warnings.warn("device='cpu' requested (floppity.py:30 default) but this implementation only runs on CUDA; "
                          "training on the current CUDA device.")
        self._round = max(self._data_round_index)
        start_idx = int(discard_prior_samples and self._round > 0)
        theta, x, masks = self.get_simulations(start_idx)
        n = theta.shape[0]
        rank, world = fdist.rank_world()
        seed = fdist.shared_seed(seed, dev)
        gen = torch.Generator().manual_seed(seed)
        n_train = int((1 - validation_fraction) * n)
        perm = torch.randperm(n, generator=gen)
        train_idx, val_idx = perm[:n_train], perm[n_train:]

        if self._neural_net is None or retrain_from_scratch:
            self._neural_net = self._build_neural_net(theta[train_idx], x[train_idx]).to(dev)
            if world > 1:
                fdist.broadcast_(self._neural_net._flat.data)
                if getattr(self._neural_net, "embedding", None) is not None and self._neural_net.embedding.param_count > 0:
                    fdist.broadcast_(self._neural_net.embedding._flat.data)
        net = self._neural_net
        net.train()

        # resident dataset: this rank's shard of the split, x standardised once
        my_train = fdist.shard_indices(train_idx) if world > 1 else train_idx
        my_val = fdist.shard_indices(val_idx) if world > 1 else val_idx
        used = torch.cat([my_train, my_val])
        if world == 1:
            # the whole dataset goes up in one copy (asynchronous when the caller's tensors are pinned) and the split is
            # gathered on the device: no host-side gather of a multi-GB array
            used_dev = used.to(dev)
            theta_dev = theta.to(dev, non_blocking=True)[used_dev].contiguous()
            x_dev = net._standardize(x.to(dev, non_blocking=True)[used_dev])
            masks_dev = masks.reshape(-1).to(dev, non_blocking=True)[used_dev].contiguous()
        else:                    # data-parallel: only this rank's shard crosses the bus, gathered into pinned staging buffers
            theta_dev = self._stage_rows(theta, used, "theta").to(dev, non_blocking=True).contiguous()
            x_dev = net._standardize(self._stage_rows(x, used, "x").to(dev, non_blocking=True))
            masks_dev = masks[used].reshape(-1).to(dev).contiguous()
        nt, nv = my_train.numel(), my_val.numel()
        B = int(min(training_batch_size, nt))
        Bv = int(min(training_batch_size, nv)) if nv > 0 else 0
        atomic = not (self._round == 0 or force_first_round_loss)
        if atomic and self._prior is None:
            raise ValueError("the atomic loss needs a prior")

        eng = _TrainEngine(net, self._prior, B, atomic, num_atoms, use_combined_loss, learning_rate, clip_max_norm,
                           use_graph=use_cuda_graph)
        eng_val = eng if Bv == B else (_TrainEngine(net, self._prior, Bv, atomic, num_atoms, use_combined_loss,
                                                    learning_rate, clip_max_norm) if Bv > 0 else None)
        self._engine = eng
        epoch, self._val_log_prob = 0, float("-inf")
        tr_ix = torch.arange(nt, device=dev)
        va_ix = torch.arange(nt, nt + nv, device=dev)

        while epoch <= max_num_epochs and not self._converged(epoch, stop_after_epochs):
            t0 = time.time()
            net.train()
            order = tr_ix[torch.randperm(nt, device=dev)]
            nb = nt // B
            eng.loss_sum.zero_()
            for b in range(nb):
                eng.step(order[b * B:(b + 1) * B], theta_dev, x_dev, masks_dev)
            train_sum = float(eng.loss_sum.item())          # the ONE host sync of the epoch
            eng.check_status()
            epoch += 1
            self._summary["training_log_probs"].append(-train_sum / max(nb * B, 1))
            net.eval()
            val_sum, val_cnt = 0.0, 0
            if eng_val is not None:
                eng_val.loss_sum.zero_()
                nvb = nv // Bv
                for b in range(nvb):
                    eng_val.eval_batch(va_ix[b * Bv:(b + 1) * Bv], theta_dev, x_dev, masks_dev)
                val_sum, val_cnt = float(eng_val.loss_sum.item()), nvb * Bv
            val_sum, val_cnt = fdist.agree_sum(val_sum, float(val_cnt), dev)
            self._val_log_prob = -val_sum / max(val_cnt, 1.0)
            self._summary["validation_log_probs"].append(self._val_log_prob)
            self._summary["epoch_durations_sec"].append(time.time() - t0)
        self._summary["epochs_trained"].append(epoch)
        self._summary["best_validation_log_prob"].append(self._best_val_log_prob)
        if show_train_summary:
            print(f"Neural network successfully converged after {epoch} epochs. "
                  f"Best validation log-prob {self._best_val_log_prob:.4f}")
        net.zero_grad(set_to_none=True)
        self._engine = None
        return copy.deepcopy(net)

    def build_posterior(self, density_estimator: Optional[NSFFlow] = None, prior: Optional[BoxUniform] = None,
                        sample_with: str = "rejection", **kwargs) -> "DirectPosterior":
        net = density_estimator if density_estimator is not None else self._neural_net
        if net is None:
            raise ValueError("train() must be called before build_posterior()")
        return DirectPosterior(net, prior if prior is not None else self._prior, device=self._device)


# ---------------------------------------------------------------------------------------------------------
class DirectPosterior:
    """``sbi.inference.posteriors.DirectPosterior``: rejection sampling of the flow inside the prior box and
    log-prob with leakage correction (floppity.py:225,252; floppityFUN.py:35,61,73)."""

    def __init__(self, posterior_estimator: NSFFlow, prior: BoxUniform, max_sampling_batch_size: int = 1 << 20,
                 device: Optional[str] = None, x_shape=None):
        self.posterior_estimator = posterior_estimator
        self.prior = prior
        self.max_sampling_batch_size = int(max_sampling_batch_size)
        self._x: Optional[Tensor] = None
        # Drop-in fidelity: FlopPITy's default is -device cpu (floppity.py:30) and its helpers call .numpy() on what
        # log_prob returns (floppityFUN.py:37-39).  The work always runs on CUDA; a posterior built for 'cpu' hands its
        # RESULTS back on the CPU, any other / no device keeps them on the GPU.
        self._return_device = torch.device("cpu") if (device is not None and str(device) == "cpu") else None
        self._leakage_cache = {}
        self._purpose = "direct posterior (B200 NSF)"

    @property
    def default_x(self):
        return self._x

    def set_default_x(self, x):
        x = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).detach().to(torch.float32).reshape(1, -1)
        if x.shape[1] != self.posterior_estimator.C:
            raise ValueError(f"x has {x.shape[1]} features, the estimator was trained on {self.posterior_estimator.C}")
        self._x = x
        return self

    def _x_or_default(self, x):
        if x is None:
            if self._x is None:
                raise ValueError("Context `x` needed when a default has not been set (set_default_x).")
            return self._x
        x = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).detach().to(torch.float32).reshape(1, -1)
        return x

    def _box(self, dev):
        return self.prior.low.to(dev).contiguous(), self.prior.high.to(dev).contiguous()

    @torch.no_grad()
    def _draw(self, n: int, ctx_std: Tensor, noise: Optional[Tensor] = None):
        """n raw flow draws + in-support mask (uint8), all on the device."""
        net = self.posterior_estimator
        dev = net._flat.device
        z = torch.randn(n, net.D, device=dev, dtype=torch.float32) if noise is None else noise
        cand = net.inverse_rows(z, ctx_std)
        inside = torch.empty(n, dtype=torch.uint8, device=dev)
        low, high = self._box(dev)
        _lib.check(_lib.lib().fp_within_box(_lib.ptr(cand), _lib.ptr(low), _lib.ptr(high), _lib.ptr(inside), n, net.D,
                                            _lib.stream()), "fp_within_box")
        return cand, inside

    @torch.no_grad()
    def _accept_reject(self, num_samples: int, x: Tensor, max_sampling_batch_size: int):
        net = self.posterior_estimator
        net.eval()
        ctx_std = net._standardize(x)
        dev = net._flat.device
        out = torch.empty(num_samples, net.D, dtype=torch.float32, device=dev)
        counter = torch.zeros(1, dtype=torch.int64, device=dev)        # accepted so far (device side)
        status = torch.zeros(4, dtype=torch.int32, device=dev)
        low, high = self._box(dev)
        num_remaining, num_total, num_accepted = num_samples, 0, 0
        bs = min(num_samples, max_sampling_batch_size)
        acceptance_rate = 1.0
        warned = False
        while num_remaining > 0:
            z = torch.randn(bs, net.D, device=dev, dtype=torch.float32)
            cand = net.inverse_rows(z, ctx_std, status)
            # in-support rows are appended to `out` on the device; the only host round trip per batch is the count
            _lib.check(_lib.lib().fp_compact_inside(_lib.ptr(cand), _lib.ptr(low), _lib.ptr(high), _lib.ptr(out),
                                                    _lib.ptr(counter), bs, net.D, num_samples, _lib.stream()),
                       "fp_compact_inside")
            num_accepted = int(counter.item())
            st = status.tolist()                      # (the count above already synchronised)
            _check_fp16_range(st)
            if st[0]:
                raise AssertionError(f"the flow produced {st[0]} non-finite samples ({st[1]} negative spline discriminants: "
                                     "upstream's `assert (discriminant >= 0).all()`)")
            num_total += bs
            num_remaining = num_samples - num_accepted
            acceptance_rate = num_accepted / num_total
            if not warned and num_total > 1000 and acceptance_rate < 0.01:
                warnings.warn(f"Only {acceptance_rate:.3%} proposal samples are accepted; the posterior leaks out of the prior.")
                warned = True
            # upstream over-draws the remainder by 1.5x; the draws are i.i.d., so the factor only trades wasted draws
            # against an extra (small) round: 1.1x once the remainder is large enough for the binomial noise not to matter
            over = 1.1 if num_remaining > 20_000 else 1.5
            bs = min(max_sampling_batch_size, max(int(over * num_remaining / max(acceptance_rate, 1e-12)), 100))
        return out, acceptance_rate

    def sample(self, sample_shape=torch.Size(), x=None, max_sampling_batch_size: Optional[int] = None,
               sample_with=None, show_progress_bars: bool = False) -> Tensor:
        """``posterior.sample((n,))`` (floppity.py:252).  Returns a tensor on the estimator's device, like upstream."""
        num_samples = int(torch.Size(sample_shape).numel())
        x = self._x_or_default(x)
        mx = self.max_sampling_batch_size if max_sampling_batch_size is None else int(max_sampling_batch_size)
        samples, _ = self._accept_reject(num_samples, x, mx)
        samples = samples.reshape(*torch.Size(sample_shape), -1)
        return samples if self._return_device is None else samples.to(self._return_device)

    def sample_sharded(self, num_samples: int, x=None, gather: bool = False) -> Tensor:
        """Embarrassingly parallel sampling: this rank draws its contiguous share of ``num_samples``."""
        lo, hi = fdist.shard_rows(num_samples)
        local = self.sample((hi - lo,), x=x)
        return fdist.gather_rows(local) if gather else local

    @torch.no_grad()
    def leakage_correction(self, x: Tensor, num_rejection_samples: int = 10_000, force_update: bool = False) -> float:
        key = x.detach().cpu().numpy().tobytes()
        if force_update or key not in self._leakage_cache:
            _, acc = self._accept_reject(num_rejection_samples, x, self.max_sampling_batch_size)
            self._leakage_cache[key] = float(acc)
        return self._leakage_cache[key]

    def log_prob(self, theta, x=None, norm_posterior: bool = True, track_gradients: bool = False,
                 leakage_correction_params: Optional[dict] = None) -> Tensor:
        """``posterior.log_prob(theta, x)`` (floppityFUN.py:35,61,73): flow log-prob with ONE broadcast context row,
        -inf outside the prior support, minus log(acceptance) when ``norm_posterior``."""
        net = self.posterior_estimator
        net.eval()
        dev = net._flat.device
        theta = torch.as_tensor(np.asarray(theta) if not torch.is_tensor(theta) else theta)
        theta = theta.to(dev, torch.float32).reshape(-1, net.D).contiguous()
        x = self._x_or_default(x)
        with torch.set_grad_enabled(track_gradients):
            lp = net.log_prob(theta, x.to(dev))
            getattr(net, "flow", net).check_last_status()      # saturated 3xFP16 operands never pass silently
            inside = torch.empty(theta.shape[0], dtype=torch.uint8, device=dev)
            low, high = self._box(dev)
            _lib.check(_lib.lib().fp_within_box(_lib.ptr(theta.detach()), _lib.ptr(low), _lib.ptr(high), _lib.ptr(inside),
                                                theta.shape[0], net.D, _lib.stream()), "fp_within_box")
            lp = torch.where(inside.bool(), lp, torch.full_like(lp, -float("inf")))
            if norm_posterior:
                acc = self.leakage_correction(x, **(leakage_correction_params or {}))
                lp = lp - float(np.log(acc))
        return lp if self._return_device is None else lp.to(self._return_device)

    def log_prob_sharded(self, theta, x=None, norm_posterior: bool = False, gather: bool = False) -> Tensor:
        theta = torch.as_tensor(theta)
        lo, hi = fdist.shard_rows(theta.shape[0])
        local = self.log_prob(theta[lo:hi], x=x, norm_posterior=norm_posterior)
        return fdist.gather_rows(local) if gather else local

    def map(self, x=None, num_iter: int = 100, num_to_optimize: int = 100, learning_rate: float = 0.01,
            init_method="posterior", num_init_samples: int = 1000, save_best_every: int = 10,
            show_progress_bars: bool = False, force_update: bool = False) -> Tensor:
        """``posterior.map`` (modules.py:173-174): gradient ascent on log_prob from the best posterior draws.
        Uses the kernels' input-gradient path (d log q / d theta)."""
        x = self._x_or_default(x)
        dev = self.posterior_estimator._flat.device
        init = (self.sample((num_init_samples,), x=x) if isinstance(init_method, str) else torch.as_tensor(init_method)).to(dev)
        lp0 = self.log_prob(init, x=x, norm_posterior=False).to(dev)
        top = init[torch.argsort(lp0, descending=True)[:num_to_optimize]].clone()
        low, high = self._box(dev)
        theta = top.clone().requires_grad_(True)
        opt = torch.optim.Adam([theta], lr=learning_rate)
        was = self.posterior_estimator._flat.requires_grad
        self.posterior_estimator._flat.requires_grad_(False)
        best, best_lp = top[0].clone(), lp0.max()
        try:
            for it in range(num_iter):
                opt.zero_grad()
                lp = self.posterior_estimator.log_prob(theta, x.to(dev))
                (-lp.sum()).backward()
                opt.step()
                with torch.no_grad():
                    theta.data = torch.minimum(torch.maximum(theta.data, low), high)
                    if it % save_best_every == 0 or it == num_iter - 1:
                        cur = self.log_prob(theta.detach(), x=x, norm_posterior=False).to(dev)
                        j = int(torch.argmax(cur))
                        if cur[j] > best_lp:
                            best, best_lp = theta.detach()[j].clone(), cur[j]
        finally:
            self.posterior_estimator._flat.requires_grad_(was)
        self._map = best if self._return_device is None else best.to(self._return_device)
        return self._map
