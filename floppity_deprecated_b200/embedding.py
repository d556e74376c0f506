"""Embedding net in front of the flow, trained jointly.

``createEmbedding`` (/root/reference/modules.py:21-45) hands ``posterior_nn(embedding_net=...)`` an
``sbi.neural_nets.embedding_nets.FCEmbedding(input_dim=nwvl, num_hiddens=256, num_layers=4, output_dim=embed_size)``
when FlopPITy runs with ``-embedding -embedding_type FC``; sbi then builds the flow on the EMBEDDED context
(``Sequential(Standardize(x), embedding)``), so the context projection of every coupling layer contracts over
``embed_size`` features instead of the raw spectrum.  This module is that path on the GPU:

* ``FCEmbedding``   -- same constructor; one flat fp32 parameter buffer; forward, dgrad and wgrad are the library's own
  GEMM kernels (tcgen05 3xTF32 where the shapes allow, exact-fp32 SIMT otherwise) with the relus folded into GEMM
  prologues / masks; upstream ``state_dict`` keys (``net.{2l}.weight / bias``).
* ``EmbeddedNSF``   -- ``Standardize -> FCEmbedding -> NSFFlow`` behind the estimator surface the rest of the package
  uses (``log_prob / sample / inverse_rows / state_dict`` with the upstream ``_embedding_net.{0,1}.*`` keys).

The CNN / multi-observation embeddings (modules.py:27-38) are not implemented and raise.
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, List, Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib


def _gemm(A, lda, a_kmajor, B, ldb, b_kmajor, C, ldc, M, N, K, flags, bias=None, msk=None, ldm=0, u=None):
    a = _lib.GemmArgs()
    a.A, a.lda, a.a_kmajor = A.data_ptr(), lda, a_kmajor
    a.B, a.ldb, a.b_kmajor = B.data_ptr(), ldb, b_kmajor
    a.C, a.ldc, a.M, a.N, a.K, a.flags, a.split_k = C.data_ptr(), ldc, M, N, K, flags, 1
    if bias is not None:
        a.bias = bias.data_ptr()
    if msk is not None:
        a.Msk, a.ldm = msk.data_ptr(), ldm
    if u is not None:
        a.U = u.data_ptr()
    return a


class FCEmbedding(nn.Module):
    """``FCEmbedding(input_dim, output_dim=20, num_layers=2, num_hiddens=50)`` (sbi 0.22 signature)."""

    def __init__(self, input_dim: int, output_dim: int = 20, num_layers: int = 2, num_hiddens: int = 50,
                 generator: Optional[torch.Generator] = None):
        super().__init__()
        if num_layers < 2:
            raise ValueError("FCEmbedding needs at least the input and the output layer")
        self.input_dim, self.output_dim = int(input_dim), int(output_dim)
        self.num_layers, self.num_hiddens = int(num_layers), int(num_hiddens)
        dims = [self.input_dim] + [self.num_hiddens] * (self.num_layers - 1) + [self.output_dim]
        self.shapes = [(dims[i + 1], dims[i]) for i in range(self.num_layers)]       # (out, in) per linear
        self.w_off, self.b_off, o = [], [], 0
        for (n_out, n_in) in self.shapes:
            self.w_off.append(o); o += (n_out * n_in + 3) // 4 * 4
            self.b_off.append(o); o += (n_out + 3) // 4 * 4
        self.param_count = o
        self._flat = nn.Parameter(torch.zeros(o, dtype=torch.float32))
        self.reset_parameters(generator)

    # ------------------------------------------------------------------ parameters
    def weight(self, l: int, flat: Optional[Tensor] = None) -> Tensor:
        flat = self._flat.detach() if flat is None else flat
        n_out, n_in = self.shapes[l]
        return flat[self.w_off[l]: self.w_off[l] + n_out * n_in].view(n_out, n_in)

    def bias(self, l: int, flat: Optional[Tensor] = None) -> Tensor:
        flat = self._flat.detach() if flat is None else flat
        return flat[self.b_off[l]: self.b_off[l] + self.shapes[l][0]]

    @torch.no_grad()
    def reset_parameters(self, generator: Optional[torch.Generator] = None) -> None:
        """torch default nn.Linear init (U(+-1/sqrt(fan_in)) for weight and bias)."""
        self._flat.zero_()
        dev = self._flat.device
        for l, (n_out, n_in) in enumerate(self.shapes):
            b = 1.0 / math.sqrt(n_in)
            self.weight(l).copy_(((torch.rand(n_out, n_in, generator=generator) * 2 - 1) * b).to(dev))
            self.bias(l).copy_(((torch.rand(n_out, generator=generator) * 2 - 1) * b).to(dev))

    def unpack(self, flat: Tensor, prefix: str = "") -> Dict[str, Tensor]:
        sd = {}
        for l in range(self.num_layers):
            sd[f"{prefix}net.{2 * l}.weight"] = self.weight(l, flat).clone()
            sd[f"{prefix}net.{2 * l}.bias"] = self.bias(l, flat).clone()
        return sd

    def state_dict(self, *args, destination=None, prefix: str = "", keep_vars: bool = False):
        sd = destination if destination is not None else {}
        sd.update(self.unpack(self._flat.detach(), prefix))
        return sd

    @torch.no_grad()
    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        for l in range(self.num_layers):
            for name, dst in (("weight", self.weight(l)), ("bias", self.bias(l))):
                key = f"net.{2 * l}.{name}"
                if key in state_dict:
                    dst.copy_(state_dict[key].detach().to(dst.device, torch.float32))
                elif strict:
                    raise KeyError(f"missing key {key}")
        return torch.nn.modules.module._IncompatibleKeys([], [])

    # ------------------------------------------------------------------ kernels
    def _split(self):
        """TF32 hi / lo copies of the weights for the tcgen05 forward GEMMs (once per optimiser step)."""
        hi, lo = torch.empty_like(self._flat), torch.empty_like(self._flat)
        _lib.check(_lib.lib().fp_split_tf32(_lib.ptr(self._flat.detach()), _lib.ptr(hi), _lib.ptr(lo), self.param_count,
                                            _lib.stream()), "fp_split_tf32")
        return hi, lo

    def forward_raw(self, x: Tensor, keep: bool):
        """x [R, input_dim] (standardised) -> (e [R, output_dim], saved pre-activations z_1..z_L when ``keep``)."""
        L, s = _lib.lib(), _lib.stream()
        if not x.is_cuda:
            raise RuntimeError("FCEmbedding runs on CUDA only; there is no CPU fallback")
        R = x.shape[0]
        flat = self._flat.detach()
        hi, lo = self._split()
        zs: List[Tensor] = []
        a_in, ld_in = x.contiguous(), self.input_dim
        for l, (n_out, n_in) in enumerate(self.shapes):
            z = torch.empty(R, n_out, dtype=torch.float32, device=x.device)
            fl = _lib.G_BIAS | (_lib.G_RELU_A if l > 0 else 0)
            W = flat[self.w_off[l]:]
            g = _gemm(a_in, ld_in, 1, W, n_in, 1, z, n_out, R, n_out, n_in, fl, bias=self.bias(l, flat))
            rc = L.fp_gemm_tc(ctypes.byref(g), _lib.ptr(hi[self.w_off[l]:]), _lib.ptr(lo[self.w_off[l]:]), s)
            if rc == -3:                                   # FP_E_NOTREADY: shape / alignment outside the tcgen05 kernel
                rc = L.fp_gemm(ctypes.byref(g), s)
            _lib.check(rc, "embedding forward")
            zs.append(z)
            a_in, ld_in = z, n_out
        e = torch.empty_like(zs[-1])
        _lib.check(L.fp_relu(_lib.ptr(zs[-1]), _lib.ptr(e), e.numel(), s), "fp_relu")
        return e, (zs if keep else None)

    def backward_raw(self, x: Tensor, zs: List[Tensor], d_e: Tensor, grads: Tensor) -> None:
        """Accumulates dL/d(parameters) into ``grads`` (laid out like ``_flat``) given dL/de [R, output_dim]."""
        L, s = _lib.lib(), _lib.stream()
        R = x.shape[0]
        flat = self._flat.detach()
        dz = torch.empty_like(d_e)
        _lib.check(L.fp_relu_bwd(_lib.ptr(d_e.contiguous()), _lib.ptr(zs[-1]), _lib.ptr(dz), dz.numel(), s), "fp_relu_bwd")
        for l in range(self.num_layers - 1, -1, -1):
            n_out, n_in = self.shapes[l]
            act = zs[l - 1] if l > 0 else x
            gW, gb = grads[self.w_off[l]:], grads[self.b_off[l]:]
            # dW[n_out, n_in] += dz^T relu(act) ; db += colsum(dz)
            fl = _lib.G_ATOMIC | (_lib.G_RELU_B if l > 0 else 0)
            g = _gemm(dz, n_out, 0, act, n_in, 0, gW, n_in, n_out, n_in, R, fl | _lib.G_COLSUM_A, u=gb)
            rc = L.fp_gemm_tc_tn(ctypes.byref(g), s)
            if rc == -3:
                g.flags, g.U, g.split_k = fl, None, max(1, min(64, R // 256))
                rc = L.fp_gemm(ctypes.byref(g), s)
                gb[:n_out] += dz.sum(0)
            _lib.check(rc, "embedding wgrad")
            if l > 0:
                # d z_{l-1} = (dz W_l) * (z_{l-1} > 0)
                dprev = torch.empty(R, n_in, dtype=torch.float32, device=x.device)
                W = flat[self.w_off[l]:]
                g = _gemm(dz, n_out, 1, W, n_in, 0, dprev, n_in, R, n_in, n_out, _lib.G_RELUMASK, msk=zs[l - 1], ldm=n_in)
                _lib.check(L.fp_gemm(ctypes.byref(g), s), "embedding dgrad")
                dz = dprev

    def forward(self, x: Tensor) -> Tensor:
        """Evaluation forward (no gradient bookkeeping: joint training goes through the trainer's engine)."""
        x = torch.as_tensor(x).to(self._flat.device, torch.float32).reshape(-1, self.input_dim)
        return self.forward_raw(x, keep=False)[0]

    def extra_repr(self) -> str:
        return f"input_dim={self.input_dim}, output_dim={self.output_dim}, num_layers={self.num_layers}, num_hiddens={self.num_hiddens}"


class EmbeddedNSF(nn.Module):
    """``Flow(transform, base, Sequential(Standardize(x), FCEmbedding))`` as sbi ``build_nsf`` assembles it when an
    embedding net is given: the flow's context is the embedding of the standardised observation."""

    def __init__(self, flow, embedding: FCEmbedding, x_mean: Optional[Tensor], x_std: Optional[Tensor]):
        super().__init__()
        self.flow = flow
        self.embedding = embedding
        self.has_zscore_x = x_mean is not None
        C = embedding.input_dim
        self.register_buffer("_x_mean", torch.zeros(C) if x_mean is None else x_mean.detach().float().clone())
        self.register_buffer("_x_std", torch.ones(C) if x_std is None else x_std.detach().float().clone())

    # -- what DirectPosterior / SNPE_C touch
    @property
    def D(self): return self.flow.D
    @property
    def C(self): return self.embedding.input_dim
    @property
    def _flat(self): return self.flow._flat
    @property
    def cfg(self): return self.flow.cfg

    def _standardize(self, x: Tensor) -> Tensor:
        x = x.to(self.flow._flat.device, torch.float32).contiguous()
        if not self.has_zscore_x:
            return x
        out = torch.empty_like(x)
        _lib.check(_lib.lib().fp_standardize(_lib.ptr(x), _lib.ptr(self._x_mean), _lib.ptr(self._x_std), _lib.ptr(out),
                                             x.shape[0], x.shape[1], _lib.stream()), "fp_standardize")
        return out

    def _embed(self, context) -> Tensor:
        dev = self.flow._flat.device
        context = torch.as_tensor(context).to(dev, torch.float32).reshape(-1, self.C)
        with torch.no_grad():
            return self.embedding.forward_raw(self._standardize(context), keep=False)[0]

    def log_prob(self, inputs, context=None) -> Tensor:
        """Gradients flow to ``inputs`` (posterior.map) -- parameter gradients are the trainer's business."""
        return self.flow.log_prob(inputs, self._embed(context))

    @torch.no_grad()
    def sample(self, num_samples: int, context=None, noise: Optional[Tensor] = None) -> Tensor:
        return self.flow.sample(num_samples, self._embed(context), noise)

    @torch.no_grad()
    def inverse_rows(self, z: Tensor, ctx_std: Tensor, status: Optional[Tensor] = None) -> Tensor:
        e = self.embedding.forward_raw(ctx_std, keep=False)[0]
        return self.flow.inverse_rows(z, e, status)

    def forward(self, inputs, context=None):
        return self.log_prob(inputs, context)

    # -- upstream-compatible checkpoint keys
    def state_dict(self, *args, destination=None, prefix: str = "", keep_vars: bool = False):
        sd = destination if destination is not None else {}
        sd.update(self.flow.state_dict(prefix=prefix))
        if self.has_zscore_x:
            sd[prefix + "_embedding_net.0._mean"] = self._x_mean.detach().clone()
            sd[prefix + "_embedding_net.0._std"] = self._x_std.detach().clone()
        base = "_embedding_net.1." if self.has_zscore_x else "_embedding_net."
        sd.update(self.embedding.state_dict(prefix=prefix + base))
        return sd

    @torch.no_grad()
    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        sd = dict(state_dict)
        base = "_embedding_net.1." if self.has_zscore_x else "_embedding_net."
        emb = {k[len(base):]: v for k, v in sd.items() if k.startswith(base)}
        self.embedding.load_state_dict(emb, strict=strict)
        if self.has_zscore_x:
            self._x_mean.copy_(sd["_embedding_net.0._mean"].to(self._x_mean.device))
            self._x_std.copy_(sd["_embedding_net.0._std"].to(self._x_std.device))
        rest = {k: v for k, v in sd.items() if not k.startswith("_embedding_net.")}
        return self.flow.load_state_dict(rest, strict=strict)


class CNNEmbedding(nn.Module):
    """Placeholder so ``from ... import FCEmbedding, CNNEmbedding, PermutationInvariantEmbedding`` (modules.py:1,
    floppityFUN.py:19) keeps importing; constructing it raises (the convolutional embeddings are not implemented)."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError("CNNEmbedding (-embedding_type CNN / multi, modules.py:27-38) is not implemented; "
                                  "use -embedding_type FC or no embedding")


class PermutationInvariantEmbedding(CNNEmbedding):
    pass
