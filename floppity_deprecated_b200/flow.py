"""``build_nsf`` / ``posterior_nn`` and the flow object they return -- drop-in for the sbi 0.22 / nflows 0.14
objects FlopPITy builds at ``/root/reference/modules.py:63`` (``density_builder`` -> ``posterior_nn(model='nsf')``).

The returned :class:`NSFFlow` keeps the nflows ``Flow`` surface the reference and sbi use
(``log_prob(inputs, context)``, ``sample(num_samples, context)``, ``parameters()``, ``state_dict()`` with the
upstream key names, ``train()/eval()``, ``to()``, deepcopy / pickle) but stores every weight in ONE flat fp32
buffer and evaluates through the hand-written sm_100a kernels of ``libfloppity_b200.so``.  No CPU fallback.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Tuple

import numpy as np
import torch
from torch import Tensor, nn

from . import _lib
from ._lib import NsfConfig

_MAX_WS_BYTES = 6 << 30  # rows are chunked so one call's workspace stays below this


def _cfg_struct(d: dict) -> NsfConfig:
    return NsfConfig(**d)


class _LogProbFn(torch.autograd.Function):
    """Autograd bridge: forward/backward are the hand-written kernels; no autograd graph inside."""

    @staticmethod
    def forward(ctx, flow: "NSFFlow", flat: Tensor, theta: Tensor, ctx_std: Tensor):
        logp, ws, prepared, seed = flow._forward_raw(theta, ctx_std, training=True)
        ctx.flow, ctx.ws, ctx.prepared, ctx.seed = flow, ws, prepared, seed
        ctx.save_for_backward(flat, ctx_std)
        ctx.R, ctx.Rc = theta.shape[0], ctx_std.shape[0]
        ctx.need_theta = theta.requires_grad
        return logp

    @staticmethod
    def backward(ctx, d_logp: Tensor):
        flow = ctx.flow
        flat, ctx_std = ctx.saved_tensors
        grads = torch.zeros(flow.grad_count, dtype=torch.float32, device=flat.device)
        d_theta = torch.empty(ctx.R, flow.D, dtype=torch.float32, device=flat.device) if ctx.need_theta else None
        flow._backward_raw(ctx_std, ctx.R, ctx.Rc, d_logp.contiguous().float(), grads, d_theta, ctx.ws, ctx.prepared,
                           ctx.seed)
        flow._finish_grads(grads, ctx.prepared)
        ctx.ws = None
        return None, grads[: flow.param_count], d_theta, None


class NSFFlow(nn.Module):
    """Conditional coupling neural spline flow (RQ-spline couplings + LU-linear layers, standard-normal base).

    Stands for ``nflows.flows.Flow(CompositeTransform([AffineTransform, T x (PiecewiseRationalQuadraticCoupling
    Transform, LULinear)]), StandardNormal, Sequential(Standardize, Identity))`` as assembled by sbi ``build_nsf``.
    """

    def __init__(self, D: int, C: int, hidden_features: int = 50, num_transforms: int = 5, num_bins: int = 10,
                 num_blocks: int = 2, dropout_probability: float = 0.0, tail_bound: float = 3.0,
                 theta_shift: Optional[Tensor] = None, theta_scale: Optional[Tensor] = None,
                 x_mean: Optional[Tensor] = None, x_std: Optional[Tensor] = None,
                 z_score_theta: bool = True, z_score_x: bool = True, generator: Optional[torch.Generator] = None):
        super().__init__()
        self.cfg = dict(D=int(D), C=int(C), T=int(num_transforms), K=int(num_bins), H=int(hidden_features),
                        Bk=int(num_blocks), tail_bound=float(tail_bound), min_bin_width=1e-3, min_bin_height=1e-3,
                        min_derivative=1e-3, lu_eps=1e-3, dropout_p=float(dropout_probability))
        self.has_zscore_theta, self.has_zscore_x = bool(z_score_theta), bool(z_score_x)
        L = _lib.lib()
        c = _cfg_struct(self.cfg)
        self.param_count = int(L.fp_nsf_param_count(c))
        self.grad_count = int(L.fp_nsf_grad_count(c))
        self.prepared_count = int(L.fp_nsf_prepared_count(c))
        if self.param_count <= 0:
            raise ValueError(f"unsupported NSF configuration {self.cfg}")
        self._index = self._build_index(L, c)
        self._flat = nn.Parameter(torch.zeros(self.param_count, dtype=torch.float32))
        self.register_buffer("_theta_shift", torch.zeros(D) if theta_shift is None else theta_shift.detach().float().clone())
        self.register_buffer("_theta_scale", torch.ones(D) if theta_scale is None else theta_scale.detach().float().clone())
        self.register_buffer("_x_mean", torch.zeros(C) if x_mean is None else x_mean.detach().float().clone())
        self.register_buffer("_x_std", torch.ones(C) if x_std is None else x_std.detach().float().clone())
        self._step_seed = 0
        self.reset_parameters(generator)

    # ------------------------------------------------------------------ layout / upstream key mapping
    @property
    def D(self): return self.cfg["D"]
    @property
    def C(self): return self.cfg["C"]
    @property
    def T(self): return self.cfg["T"]
    @property
    def hidden_features(self): return self.cfg["H"]

    def _build_index(self, L, c) -> Dict[Tuple[int, int, int], Tuple[int, int]]:
        idx = {}
        for i in range(self.cfg["T"]):
            for kind in range(13):
                nblk = {_lib.P_CTX_W: self.cfg["Bk"] + 1, _lib.P_CTX_B: self.cfg["Bk"] + 1, _lib.P_W1: self.cfg["Bk"],
                        _lib.P_B1: self.cfg["Bk"], _lib.P_W2: self.cfg["Bk"], _lib.P_B2: self.cfg["Bk"]}.get(kind, 1)
                for b in range(nblk):
                    off = int(L.fp_nsf_param_offset(c, i, kind, b))
                    size = int(L.fp_nsf_param_size(c, i, kind))
                    if off < 0 or size < 0:
                        raise RuntimeError("layout query failed")
                    idx[(i, kind, b)] = (off, size)
        return idx

    def view(self, layer: int, kind: int, block: int = 0, flat: Optional[Tensor] = None) -> Tensor:
        """View of one upstream tensor inside the flat buffer."""
        off, size = self._index[(layer, kind, block)]
        flat = self._flat.detach() if flat is None else flat
        v = flat[off: off + size]
        H, D, C = self.cfg["H"], self.cfg["D"], self.cfg["C"]
        shape = {_lib.P_CTX_W: (H, C), _lib.P_W0X: (H, D), _lib.P_W1: (H, H), _lib.P_W2: (H, H),
                 _lib.P_WF: (size // H, H)}.get(kind)
        return v.view(*shape) if shape else v

    def transform_features(self, layer: int) -> Tensor:
        return torch.arange(layer % 2, self.D, 2)

    def identity_features(self, layer: int) -> Tensor:
        return torch.arange(1 - layer % 2, self.D, 2)

    @torch.no_grad()
    def reset_parameters(self, generator: Optional[torch.Generator] = None) -> None:
        """torch default ``nn.Linear`` init everywhere except each block's second linear layer (U(+-1e-3)); LULinear
        starts at the identity (nflows ResidualBlock zero_initialization / LULinear identity_init)."""
        H, D, C, Bk = self.cfg["H"], self.cfg["D"], self.cfg["C"], self.cfg["Bk"]
        dev = self._flat.device

        def uni(shape, bound):
            return (torch.rand(shape, generator=generator) * 2 - 1).mul_(bound).to(dev)

        self._flat.zero_()
        for i in range(self.T):
            idf = self.identity_features(i).to(dev)
            fan_in = len(idf) + C
            b = 1.0 / math.sqrt(fan_in)
            w0 = uni((H, fan_in), b)
            self.view(i, _lib.P_W0X)[:, idf] = w0[:, : len(idf)]
            self.view(i, _lib.P_CTX_W, 0).copy_(w0[:, len(idf):])
            self.view(i, _lib.P_CTX_B, 0).copy_(uni((H,), b))
            for k in range(Bk):
                self.view(i, _lib.P_CTX_W, 1 + k).copy_(uni((H, C), 1.0 / math.sqrt(C)))
                self.view(i, _lib.P_CTX_B, 1 + k).copy_(uni((H,), 1.0 / math.sqrt(C)))
                self.view(i, _lib.P_W1, k).copy_(uni((H, H), 1.0 / math.sqrt(H)))
                self.view(i, _lib.P_B1, k).copy_(uni((H,), 1.0 / math.sqrt(H)))
                self.view(i, _lib.P_W2, k).copy_(uni((H, H), 1e-3))
                self.view(i, _lib.P_B2, k).copy_(uni((H,), 1e-3))
            dP = self.view(i, _lib.P_BF).numel()
            self.view(i, _lib.P_WF).copy_(uni((dP, H), 1.0 / math.sqrt(H)))
            self.view(i, _lib.P_BF).copy_(uni((dP,), 1.0 / math.sqrt(H)))
            self.view(i, _lib.P_LU_DIAG).fill_(float(np.log(np.exp(1 - 1e-3) - 1)))

    # ------------------------------------------------------------------ upstream-compatible state_dict
    def _key_prefixes(self, i: int) -> Tuple[str, str]:
        base = 1 if self.has_zscore_theta else 0
        return f"_transform._transforms.{base + 2 * i}.", f"_transform._transforms.{base + 2 * i + 1}."

    def unpack(self, flat: Tensor, prefix: str = "") -> Dict[str, Tensor]:
        """Slice a flat buffer laid out like the parameters (weights, gradients, Adam moments) into tensors keyed
        and shaped like the upstream nflows/sbi ``state_dict``."""
        sd = {}
        Bk = self.cfg["Bk"]
        dev = flat.device
        for i in range(self.T):
            cp, lp = self._key_prefixes(i)
            cp, lp = prefix + cp, prefix + lp
            idf = self.identity_features(i).to(dev)
            w0 = self.view(i, _lib.P_W0X, 0, flat)[:, idf]
            sd[cp + "transform_net.initial_layer.weight"] = torch.cat([w0, self.view(i, _lib.P_CTX_W, 0, flat)], dim=1)
            sd[cp + "transform_net.initial_layer.bias"] = self.view(i, _lib.P_CTX_B, 0, flat).clone()
            for k in range(Bk):
                bp = cp + f"transform_net.blocks.{k}."
                sd[bp + "context_layer.weight"] = self.view(i, _lib.P_CTX_W, 1 + k, flat).clone()
                sd[bp + "context_layer.bias"] = self.view(i, _lib.P_CTX_B, 1 + k, flat).clone()
                sd[bp + "linear_layers.0.weight"] = self.view(i, _lib.P_W1, k, flat).clone()
                sd[bp + "linear_layers.0.bias"] = self.view(i, _lib.P_B1, k, flat).clone()
                sd[bp + "linear_layers.1.weight"] = self.view(i, _lib.P_W2, k, flat).clone()
                sd[bp + "linear_layers.1.bias"] = self.view(i, _lib.P_B2, k, flat).clone()
            sd[cp + "transform_net.final_layer.weight"] = self.view(i, _lib.P_WF, 0, flat).clone()
            sd[cp + "transform_net.final_layer.bias"] = self.view(i, _lib.P_BF, 0, flat).clone()
            sd[lp + "bias"] = self.view(i, _lib.P_LU_BIAS, 0, flat).clone()
            sd[lp + "lower_entries"] = self.view(i, _lib.P_LU_LOWER, 0, flat).clone()
            sd[lp + "upper_entries"] = self.view(i, _lib.P_LU_UPPER, 0, flat).clone()
            sd[lp + "unconstrained_upper_diag"] = self.view(i, _lib.P_LU_DIAG, 0, flat).clone()
        return sd

    def state_dict(self, *args, destination=None, prefix: str = "", keep_vars: bool = False):
        """Keys and shapes follow nflows/sbi so checkpoints interoperate with ``posteriors.pt``."""
        sd = destination if destination is not None else {}
        dev = self._flat.device
        if self.has_zscore_theta:
            sd[prefix + "_transform._transforms.0._shift"] = self._theta_shift.detach().clone()
            sd[prefix + "_transform._transforms.0._scale"] = self._theta_scale.detach().clone()
        for i in range(self.T):
            cp, _ = self._key_prefixes(i)
            sd[prefix + cp + "identity_features"] = self.identity_features(i).to(dev)
            sd[prefix + cp + "transform_features"] = self.transform_features(i).to(dev)
        sd.update(self.unpack(self._flat.detach(), prefix))
        if self.has_zscore_x:
            sd[prefix + "_embedding_net.0._mean"] = self._x_mean.detach().clone()
            sd[prefix + "_embedding_net.0._std"] = self._x_std.detach().clone()
        return sd

    @torch.no_grad()
    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        sd = dict(state_dict)
        dev = self._flat.device
        Bk, D = self.cfg["Bk"], self.D
        used = set()

        def take(key):
            if key not in sd:
                if strict:
                    raise KeyError(f"missing key {key}")
                return None
            used.add(key)
            return sd[key].detach().to(dev, torch.float32)

        if self.has_zscore_theta:
            for name, buf in (("_shift", self._theta_shift), ("_scale", self._theta_scale)):
                v = take("_transform._transforms.0." + name)
                if v is not None:
                    buf.copy_(v)
        for i in range(self.T):
            cp, lp = self._key_prefixes(i)
            for k in (cp + "identity_features", cp + "transform_features"):
                used.add(k)
            idf = self.identity_features(i).to(dev)
            w = take(cp + "transform_net.initial_layer.weight")
            if w is not None:
                w0x = self.view(i, _lib.P_W0X)
                w0x.zero_()
                w0x[:, idf] = w[:, : len(idf)]
                self.view(i, _lib.P_CTX_W, 0).copy_(w[:, len(idf):])
            for key, kind, blk in [("transform_net.initial_layer.bias", _lib.P_CTX_B, 0),
                                   ("transform_net.final_layer.weight", _lib.P_WF, 0),
                                   ("transform_net.final_layer.bias", _lib.P_BF, 0)]:
                v = take(cp + key)
                if v is not None:
                    self.view(i, kind, blk).copy_(v)
            for k in range(Bk):
                bp = cp + f"transform_net.blocks.{k}."
                for key, kind, blk in [("context_layer.weight", _lib.P_CTX_W, 1 + k), ("context_layer.bias", _lib.P_CTX_B, 1 + k),
                                       ("linear_layers.0.weight", _lib.P_W1, k), ("linear_layers.0.bias", _lib.P_B1, k),
                                       ("linear_layers.1.weight", _lib.P_W2, k), ("linear_layers.1.bias", _lib.P_B2, k)]:
                    v = take(bp + key)
                    if v is not None:
                        self.view(i, kind, blk).copy_(v)
            for key, kind in [("bias", _lib.P_LU_BIAS), ("lower_entries", _lib.P_LU_LOWER),
                              ("upper_entries", _lib.P_LU_UPPER), ("unconstrained_upper_diag", _lib.P_LU_DIAG)]:
                v = take(lp + key)
                if v is not None:
                    self.view(i, kind).copy_(v)
        if self.has_zscore_x:
            for name, buf in (("_mean", self._x_mean), ("_std", self._x_std)):
                v = take("_embedding_net.0." + name)
                if v is not None:
                    buf.copy_(v)
        extra = [k for k in sd if k not in used and not k.endswith("_log_z")]
        if strict and extra:
            raise KeyError(f"unexpected keys {extra[:4]}...")
        return torch.nn.modules.module._IncompatibleKeys([], extra)

    # ------------------------------------------------------------------ raw kernel calls
    def _zlogdet(self) -> float:
        return float(torch.log(torch.abs(self._theta_scale.double())).sum().item())

    def _prepare(self, flat: Optional[Tensor] = None) -> Tensor:
        flat = self._flat if flat is None else flat
        prepared = torch.empty(self.prepared_count, dtype=torch.float32, device=flat.device)
        c = _cfg_struct(self.cfg)
        _lib.check(_lib.lib().fp_nsf_prepare(c, _lib.ptr(flat), _lib.ptr(prepared), _lib.stream()), "fp_nsf_prepare")
        return prepared

    def _status(self) -> Tensor:
        return torch.zeros(4, dtype=torch.int32, device=self._flat.device)

    def _standardize(self, x: Tensor) -> Tensor:
        x = x.to(self._flat.device, torch.float32).contiguous()
        if not self.has_zscore_x:
            return x
        out = torch.empty_like(x)
        _lib.check(_lib.lib().fp_standardize(_lib.ptr(x), _lib.ptr(self._x_mean), _lib.ptr(self._x_std), _lib.ptr(out),
                                             x.shape[0], x.shape[1], _lib.stream()), "fp_standardize")
        return out

    def workspace_floats(self, R: int, Rc: int, training: bool) -> int:
        n = int(_lib.lib().fp_nsf_workspace_count(_cfg_struct(self.cfg), R, Rc, int(training)))
        if n < 0:
            raise RuntimeError("fp_nsf_workspace_count failed")
        return n

    def _forward_raw(self, theta: Tensor, ctx_std: Tensor, training: bool, prepared: Optional[Tensor] = None,
                     ws: Optional[Tensor] = None, status: Optional[Tensor] = None, zlogdet: Optional[float] = None):
        _lib.require_cuda()
        R, Rc = theta.shape[0], ctx_std.shape[0]
        dev = self._flat.device
        if prepared is None:
            prepared = self._prepare()
        if ws is None:
            ws = torch.empty(self.workspace_floats(R, Rc, training), dtype=torch.float32, device=dev)
        logp = torch.empty(R, dtype=torch.float32, device=dev)
        if status is None:
            status = self._status()
        self._last_status = status       # read lazily by whoever hands the result to the user (check_last_status)
        seed = 0
        if training and self.cfg["dropout_p"] > 0 and self.training:
            self._step_seed += 1
            seed = self._step_seed
        c = _cfg_struct(self.cfg)
        if not (training and self.training):
            c.dropout_p = 0.0
        _lib.check(_lib.lib().fp_nsf_logprob_forward(
            c, _lib.ptr(self._flat), _lib.ptr(prepared), _lib.ptr(theta), _lib.ptr(ctx_std), R, Rc,
            _lib.ptr(self._theta_shift), _lib.ptr(self._theta_scale), self._zlogdet() if zlogdet is None else zlogdet,
            _lib.ptr(logp), _lib.ptr(ws), int(training), seed, _lib.ptr(status), _lib.stream()), "fp_nsf_logprob_forward")
        return logp, ws, prepared, seed

    def _backward_raw(self, ctx_std, R, Rc, d_logp, grads, d_theta, ws, prepared, seed):
        c = _cfg_struct(self.cfg)
        if not self.training or seed == 0:
            c.dropout_p = 0.0
        _lib.check(_lib.lib().fp_nsf_logprob_backward(
            c, _lib.ptr(self._flat), _lib.ptr(prepared), _lib.ptr(ctx_std), R, Rc, _lib.ptr(self._theta_scale),
            _lib.ptr(d_logp), _lib.ptr(grads), _lib.ptr(d_theta), _lib.ptr(ws), seed,
            _lib.ptr(self._last_status) if getattr(self, "_last_status", None) is not None else None, _lib.stream()),
            "fp_nsf_logprob_backward")

    def _finish_grads(self, grads, prepared):
        _lib.check(_lib.lib().fp_nsf_finish_grads(_cfg_struct(self.cfg), _lib.ptr(self._flat), _lib.ptr(prepared),
                                                  _lib.ptr(grads), _lib.stream()), "fp_nsf_finish_grads")

    def check_last_status(self) -> None:
        """Raise if the last forward / inverse pass reported saturated 3xFP16 operands (status word 2): one small
        device-to-host read, so callers do it where a result is about to leave the device anyway."""
        st = getattr(self, "_last_status", None)
        if st is None:
            return
        st = st.tolist()
        if len(st) > 2 and st[2]:
            raise FloatingPointError(
                f"{st[2]} operand pieces of the fused conditioner reached the edge of fp16's range (|value| > 6e4): the "
                "3xFP16 operand format does not cover this flow's activations -- set FLOPPITY_FUSED_FP16=0 (3xTF32 "
                "operands) and check that the observations are z-scored")

    def _rows_per_chunk(self, Rc_is_one: bool) -> int:
        per_row = self.workspace_floats(1024, 1 if Rc_is_one else 1024, False) / 1024.0
        return max(1024, int(_MAX_WS_BYTES / 4 / max(per_row, 1.0)) // 1024 * 1024)

    # ------------------------------------------------------------------ nflows Flow surface
    def log_prob(self, inputs, context=None) -> Tensor:
        """``Flow.log_prob(inputs[B, D], context[B, C]) -> [B]`` (floppityFUN.py:35 via DirectPosterior; the training
        loss floppity.py:219).  A single context row is broadcast WITHOUT materialising the repeat."""
        dev = _lib.require_cuda(self._flat.device if self._flat.is_cuda else None)
        if not self._flat.is_cuda:
            raise RuntimeError("NSFFlow parameters are on the CPU; call .to('cuda') -- there is no CPU fallback")
        inputs = torch.as_tensor(inputs)
        theta = inputs.to(dev, torch.float32)
        theta = theta.reshape(-1, self.D).contiguous()
        context = torch.as_tensor(context).to(dev, torch.float32).reshape(-1, self.C)
        R, Rc = theta.shape[0], context.shape[0]
        if Rc != R and Rc != 1 and (R % Rc != 0):
            raise ValueError("Number of input items must be equal to number of context items.")
        ctx_std = self._standardize(context)
        needs_grad = torch.is_grad_enabled() and (self._flat.requires_grad or theta.requires_grad)
        if needs_grad:
            return _LogProbFn.apply(self, self._flat, theta, ctx_std)
        status = self._status()
        prepared = self._prepare()
        zl = self._zlogdet()
        chunk = self._rows_per_chunk(Rc == 1)
        if R <= chunk:
            logp, _, _, _ = self._forward_raw(theta, ctx_std, False, prepared, None, status, zl)
            return logp
        outs = []
        A = R // Rc if Rc > 1 else 1
        chunk = max(A, chunk // A * A)
        for r0 in range(0, R, chunk):
            r1 = min(R, r0 + chunk)
            cs = ctx_std if Rc == 1 else ctx_std[r0 // A: r1 // A]
            outs.append(self._forward_raw(theta[r0:r1], cs, False, prepared, None, status, zl)[0])
        return torch.cat(outs)

    @torch.no_grad()
    def sample(self, num_samples: int, context=None, noise: Optional[Tensor] = None) -> Tensor:
        """``Flow.sample(num_samples, context[M, C]) -> [M, num_samples, D]`` (floppity.py:252 via DirectPosterior).
        ``noise`` ([M*num_samples, D]) replaces the base draw, so parity tests never depend on RNG streams."""
        dev = _lib.require_cuda()
        context = torch.as_tensor(context).to(dev, torch.float32).reshape(-1, self.C)
        M = context.shape[0]
        R = M * int(num_samples)
        ctx_std = self._standardize(context)
        z = torch.randn(R, self.D, device=dev, dtype=torch.float32) if noise is None else \
            torch.as_tensor(noise).to(dev, torch.float32).reshape(R, self.D).contiguous()
        out = self.inverse_rows(z, ctx_std)
        return out.reshape(M, int(num_samples), self.D)

    @torch.no_grad()
    def inverse_rows(self, z: Tensor, ctx_std: Tensor, status: Optional[Tensor] = None) -> Tensor:
        """Inverse pass on base-noise rows ``z [R, D]`` with standardised context ``[Rc, C]`` (Rc | R)."""
        dev = self._flat.device
        R, Rc = z.shape[0], ctx_std.shape[0]
        out = torch.empty(R, self.D, dtype=torch.float32, device=dev)
        if R == 0:
            return out
        prepared = self._prepare()
        status = self._status() if status is None else status
        c = _cfg_struct(self.cfg)
        A = R // Rc
        chunk = self._rows_per_chunk(Rc == 1)
        chunk = max(A, chunk // A * A) if Rc > 1 else chunk
        for r0 in range(0, R, chunk):
            r1 = min(R, r0 + chunk)
            cs = ctx_std if Rc == 1 else ctx_std[r0 // A: r1 // A]
            ws = torch.empty(self.workspace_floats(r1 - r0, cs.shape[0], False), dtype=torch.float32, device=dev)
            _lib.check(_lib.lib().fp_nsf_sample(
                c, _lib.ptr(self._flat), _lib.ptr(prepared), _lib.ptr(z[r0:r1]), _lib.ptr(cs), r1 - r0, cs.shape[0],
                _lib.ptr(self._theta_shift), _lib.ptr(self._theta_scale), _lib.ptr(out[r0:r1]), _lib.ptr(ws),
                _lib.ptr(status), _lib.stream()), "fp_nsf_sample")
        self._last_status = status
        return out

    def forward(self, inputs, context=None):
        return self.log_prob(inputs, context)

    def extra_repr(self) -> str:
        return ", ".join(f"{k}={v}" for k, v in self.cfg.items())


# ---------------------------------------------------------------------------------------------------------
# sbi.utils.sbiutils z-scoring + sbi.neural_nets.flow.build_nsf + sbi.utils.get_nn_models.posterior_nn
# ---------------------------------------------------------------------------------------------------------
def _valid_rows(t: Tensor) -> Tensor:
    flat = t.reshape(t.shape[0], -1)
    return ~(torch.isnan(flat).any(dim=1) | torch.isinf(flat).any(dim=1))


def _zscore_stats(batch: Tensor, min_std: float) -> Tuple[Tensor, Tensor]:
    ok = _valid_rows(batch)
    mean = torch.mean(batch[ok], dim=0)
    std = torch.std(batch[ok], dim=0)
    std[std < min_std] = min_std
    return mean, std


def _zscore_on(flag) -> bool:
    if flag in (None, False, "none"):
        return False
    if flag in (True, "independent"):
        return True
    raise NotImplementedError(f"z_score={flag!r}: only 'independent' / 'none' are on the FlopPITy path")


def build_nsf(batch_x: Tensor, batch_y: Tensor, z_score_x="independent", z_score_y="independent",
              hidden_features: int = 50, num_transforms: int = 5, num_bins: int = 10,
              embedding_net: Optional[nn.Module] = None, tail_bound: float = 3.0,
              hidden_layers_spline_context: int = 1, num_blocks: int = 2, dropout_probability: float = 0.0,
              use_batch_norm: bool = False, generator: Optional[torch.Generator] = None, **kwargs) -> NSFFlow:
    """sbi ``build_nsf`` (sbi naming: ``batch_x`` is theta, ``batch_y`` the spectrum).  Statistics for the two
    z-scorings come from the given batches exactly as upstream (unbiased std, clamp 1e-14 / 1e-7)."""
    from .embedding import CNNEmbedding, EmbeddedNSF, FCEmbedding, FrozenEmbedding
    emb = None
    if embedding_net is not None and not isinstance(embedding_net, nn.Identity):
        if isinstance(embedding_net, (FCEmbedding, CNNEmbedding)):        # -embedding_type FC / CNN (modules.py:25-33)
            emb = embedding_net
        elif isinstance(embedding_net, nn.Module) and not any(p.requires_grad and p.numel() for p in embedding_net.parameters()):
            # no registered trainable parameters: a fixed feature map (the reference's multiNet, floppityFUN.py:93-134)
            _y = torch.as_tensor(batch_y, dtype=torch.float32)
            with torch.no_grad():
                probe = embedding_net(_y[:2].reshape(2, -1))
            emb = FrozenEmbedding(embedding_net, _y[0].numel(), probe.reshape(2, -1).shape[1])
        else:
            raise NotImplementedError("embedding nets with trainable parameters must be this package's FCEmbedding / CNNEmbedding "
                                      "(modules.py:25-33): their backward passes are hand-written kernels, not autograd graphs")
    if use_batch_norm:
        raise NotImplementedError("use_batch_norm=True is not on the FlopPITy path")
    batch_x = torch.as_tensor(batch_x, dtype=torch.float32)
    batch_y = torch.as_tensor(batch_y, dtype=torch.float32)
    D, C = batch_x[0].numel(), batch_y[0].numel()
    if D < 2:
        raise NotImplementedError("1-D theta uses a different conditioner upstream; not on the FlopPITy path")
    zx, zy = _zscore_on(z_score_x), _zscore_on(z_score_y)
    shift = scale = mean = std = None
    if zx:
        t_mean, t_std = _zscore_stats(batch_x.reshape(batch_x.shape[0], -1).cpu(), 1e-14)
        shift, scale = -t_mean / t_std, 1 / t_std
    if zy:
        mean, std = _zscore_stats(batch_y.reshape(batch_y.shape[0], -1).cpu(), 1e-7)
    if emb is not None:
        # sbi: Flow(..., embedding_net=Sequential(Standardize(y), embedding)); the flow's context is the embedding output
        if emb.input_dim != C:
            raise ValueError(f"embedding net expects {emb.input_dim} input features, the observations have {C}")
        flow = NSFFlow(D, emb.output_dim, hidden_features=hidden_features, num_transforms=num_transforms,
                       num_bins=num_bins, num_blocks=num_blocks, dropout_probability=dropout_probability,
                       tail_bound=tail_bound, theta_shift=shift, theta_scale=scale, z_score_theta=zx, z_score_x=False,
                       generator=generator)
        import copy
        return EmbeddedNSF(flow, copy.deepcopy(emb), mean, std)
    return NSFFlow(D, C, hidden_features=hidden_features, num_transforms=num_transforms, num_bins=num_bins,
                   num_blocks=num_blocks, dropout_probability=dropout_probability, tail_bound=tail_bound,
                   theta_shift=shift, theta_scale=scale, x_mean=mean, x_std=std, z_score_theta=zx, z_score_x=zy,
                   generator=generator)


class _NSFBuilder:
    """Picklable ``build_fn(batch_theta, batch_x)`` (upstream returns a closure and has to drop it when the
    inference object is pickled at floppity.py:247; a class keeps ``-resume`` able to retrain from scratch)."""

    def __init__(self, **kw):
        self.kw = kw

    def __call__(self, batch_theta, batch_x) -> NSFFlow:
        kw = dict(self.kw)
        return build_nsf(batch_x=batch_theta, batch_y=batch_x, z_score_x=kw.pop("z_score_theta"),
                         z_score_y=kw.pop("z_score_x"), **kw)


def posterior_nn(model: str, z_score_theta="independent", z_score_x="independent", hidden_features: int = 50,
                 num_transforms: int = 5, num_bins: int = 10, embedding_net: Optional[nn.Module] = None,
                 num_components: int = 10, **kwargs) -> Callable:
    """sbi ``posterior_nn`` as called at modules.py:63: returns ``build_fn(batch_theta, batch_x)``."""
    if model != "nsf":
        raise NotImplementedError("only model='nsf' is on the FlopPITy hot path (floppity.py:43 default)")
    return _NSFBuilder(z_score_theta=z_score_theta, z_score_x=z_score_x, hidden_features=hidden_features,
                       num_transforms=num_transforms, num_bins=num_bins, embedding_net=embedding_net, **kwargs)
