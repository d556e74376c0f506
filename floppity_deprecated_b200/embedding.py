"""Embedding nets in front of the flow, trained jointly.

``createEmbedding`` (/root/reference/modules.py:21-45) hands ``posterior_nn(embedding_net=...)`` one of

* ``FCEmbedding(input_dim=nwvl, num_hiddens=256, num_layers=4, output_dim=embed_size)``      (``-embedding_type FC``),
* ``CNNEmbedding(input_shape=(nwvl,), out_channels_per_layer, num_conv_layers, kernel_size, num_linear_layers,
  num_linear_units, output_dim)``                                     (``-embedding_type CNN``, the default type),
* ``multiNet`` (/root/reference/floppityFUN.py:93-134): one CNNEmbedding per observation on column slices, concatenated
  (``-embedding_type multi``).  The reference keeps those sub-nets in a plain Python list, so their parameters are not
  registered and never trained: it is a FIXED random feature map, and is treated as exactly that here
  (``FrozenEmbedding``).

sbi then builds the flow on the EMBEDDED context (``Sequential(Standardize(x), embedding)``), so the context projection
of every coupling layer contracts over ``embed_size`` features instead of the raw spectrum.  This module is that path on
the GPU:

* ``FCEmbedding``  -- same constructor; one flat fp32 parameter buffer; forward, dgrad and wgrad are the library's own GEMM
  kernels (tcgen05 3xTF32 where the shapes allow, exact-fp32 SIMT otherwise) with the relus folded into GEMM prologues /
  masks; upstream ``state_dict`` keys (``net.{2l}.weight / bias``).
* ``CNNEmbedding`` -- same constructor (1-D spectra); ``num_conv_layers x [Conv1d(padding 1) -> ReLU -> MaxPool1d]`` as one
  fused forward kernel and three backward kernels per layer (``csrc/embed_conv.cu``), followed by the FC core; upstream keys
  (``cnn_subnet.{3i}.weight / bias``, ``linear_subnet.net.{2l}.weight / bias``).
* ``EmbeddedNSF``  -- ``Standardize -> embedding -> NSFFlow`` behind the estimator surface the rest of the package uses
  (``log_prob / sample / inverse_rows / state_dict`` with the upstream ``_embedding_net.{0,1}.*`` keys).
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, List, Optional, Sequence

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


class _FCCore:
    """The arithmetic of sbi ``FCEmbedding`` on a slice ``[base, base + count)`` of somebody's flat parameter buffer:
    ``Linear -> ReLU`` x num_layers (the last ReLU included, as upstream)."""

    def __init__(self, input_dim: int, output_dim: int, num_layers: int, num_hiddens: int, base: int = 0):
        if num_layers < 2:
            raise ValueError("FCEmbedding needs at least the input and the output layer")
        self.input_dim, self.output_dim = int(input_dim), int(output_dim)
        self.num_layers, self.num_hiddens = int(num_layers), int(num_hiddens)
        dims = [self.input_dim] + [self.num_hiddens] * (self.num_layers - 1) + [self.output_dim]
        self.shapes = [(dims[i + 1], dims[i]) for i in range(self.num_layers)]       # (out, in) per linear
        self.base = int(base)
        self.w_off, self.b_off, o = [], [], self.base
        for (n_out, n_in) in self.shapes:
            self.w_off.append(o); o += (n_out * n_in + 3) // 4 * 4
            self.b_off.append(o); o += (n_out + 3) // 4 * 4
        self.count = o - self.base

    def weight(self, l: int, flat: Tensor) -> Tensor:
        n_out, n_in = self.shapes[l]
        return flat[self.w_off[l]: self.w_off[l] + n_out * n_in].view(n_out, n_in)

    def bias(self, l: int, flat: Tensor) -> Tensor:
        return flat[self.b_off[l]: self.b_off[l] + self.shapes[l][0]]

    @torch.no_grad()
    def reset(self, flat: Tensor, generator: Optional[torch.Generator] = None) -> None:
        """torch default nn.Linear init (U(+-1/sqrt(fan_in)) for weight and bias)."""
        dev = flat.device
        for l, (n_out, n_in) in enumerate(self.shapes):
            b = 1.0 / math.sqrt(n_in)
            self.weight(l, flat).copy_(((torch.rand(n_out, n_in, generator=generator) * 2 - 1) * b).to(dev))
            self.bias(l, flat).copy_(((torch.rand(n_out, generator=generator) * 2 - 1) * b).to(dev))

    def unpack(self, flat: Tensor, prefix: str = "") -> Dict[str, Tensor]:
        sd = {}
        for l in range(self.num_layers):
            sd[f"{prefix}net.{2 * l}.weight"] = self.weight(l, flat).clone()
            sd[f"{prefix}net.{2 * l}.bias"] = self.bias(l, flat).clone()
        return sd

    @torch.no_grad()
    def load(self, flat: Tensor, state_dict, prefix: str = "", strict: bool = True) -> None:
        for l in range(self.num_layers):
            for name, dst in (("weight", self.weight(l, flat)), ("bias", self.bias(l, flat))):
                key = f"{prefix}net.{2 * l}.{name}"
                if key in state_dict:
                    dst.copy_(state_dict[key].detach().to(dst.device, torch.float32))
                elif strict:
                    raise KeyError(f"missing key {key}")

    def forward_raw(self, flat: Tensor, x: Tensor, keep: bool):
        """x [R, input_dim] -> (e [R, output_dim], saved pre-activations z_1..z_L when ``keep``)."""
        L, s = _lib.lib(), _lib.stream()
        if not x.is_cuda:
            raise RuntimeError("the embedding nets run on CUDA only; there is no CPU fallback")
        R = x.shape[0]
        # TF32 hi / lo copies of the weights for the tcgen05 forward GEMMs (once per call = once per optimiser step)
        hi, lo = torch.empty_like(flat), torch.empty_like(flat)
        _lib.check(L.fp_split_tf32(_lib.ptr(flat), _lib.ptr(hi), _lib.ptr(lo), flat.numel(), s), "fp_split_tf32")
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

    def backward_raw(self, flat: Tensor, x: Tensor, zs: List[Tensor], d_e: Tensor, grads: Tensor, want_dx: bool = False):
        """Accumulates dL/d(parameters) into ``grads`` (laid out like the flat buffer) given dL/de [R, output_dim];
        returns dL/dx [R, input_dim] when ``want_dx`` (a convolutional stack in front needs it)."""
        L, s = _lib.lib(), _lib.stream()
        R = x.shape[0]
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
            if l > 0 or want_dx:
                # d z_{l-1} = (dz W_l) * (z_{l-1} > 0)      (l == 0: the plain input gradient)
                dprev = torch.empty(R, n_in, dtype=torch.float32, device=x.device)
                W = flat[self.w_off[l]:]
                if l > 0:
                    g = _gemm(dz, n_out, 1, W, n_in, 0, dprev, n_in, R, n_in, n_out, _lib.G_RELUMASK, msk=zs[l - 1], ldm=n_in)
                else:
                    g = _gemm(dz, n_out, 1, W, n_in, 0, dprev, n_in, R, n_in, n_out, 0)
                _lib.check(L.fp_gemm(ctypes.byref(g), s), "embedding dgrad")
                dz = dprev
        return dz if want_dx else None


class FCEmbedding(nn.Module):
    """``FCEmbedding(input_dim, output_dim=20, num_layers=2, num_hiddens=50)`` (sbi 0.22 signature)."""

    def __init__(self, input_dim: int, output_dim: int = 20, num_layers: int = 2, num_hiddens: int = 50,
                 generator: Optional[torch.Generator] = None):
        super().__init__()
        self.core = _FCCore(input_dim, output_dim, num_layers, num_hiddens)
        self.input_dim, self.output_dim = self.core.input_dim, self.core.output_dim
        self.num_layers, self.num_hiddens = self.core.num_layers, self.core.num_hiddens
        self.shapes, self.w_off, self.b_off = self.core.shapes, self.core.w_off, self.core.b_off
        self.param_count = self.core.count
        self._flat = nn.Parameter(torch.zeros(self.param_count, dtype=torch.float32))
        self.reset_parameters(generator)

    def weight(self, l: int, flat: Optional[Tensor] = None) -> Tensor:
        return self.core.weight(l, self._flat.detach() if flat is None else flat)

    def bias(self, l: int, flat: Optional[Tensor] = None) -> Tensor:
        return self.core.bias(l, self._flat.detach() if flat is None else flat)

    @torch.no_grad()
    def reset_parameters(self, generator: Optional[torch.Generator] = None) -> None:
        self._flat.zero_()
        self.core.reset(self._flat, generator)

    def unpack(self, flat: Tensor, prefix: str = "") -> Dict[str, Tensor]:
        return self.core.unpack(flat, prefix)

    def state_dict(self, *args, destination=None, prefix: str = "", keep_vars: bool = False):
        sd = destination if destination is not None else {}
        sd.update(self.unpack(self._flat.detach(), prefix))
        return sd

    @torch.no_grad()
    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        self.core.load(self._flat, state_dict, "", strict)
        return torch.nn.modules.module._IncompatibleKeys([], [])

    def forward_raw(self, x: Tensor, keep: bool):
        return self.core.forward_raw(self._flat.detach(), x, keep)

    def backward_raw(self, x: Tensor, zs: List[Tensor], d_e: Tensor, grads: Tensor) -> None:
        self.core.backward_raw(self._flat.detach(), x, zs, d_e, grads)

    def forward(self, x: Tensor) -> Tensor:
        """Evaluation forward (no gradient bookkeeping: joint training goes through the trainer's engine)."""
        x = torch.as_tensor(x).to(self._flat.device, torch.float32).reshape(-1, self.input_dim)
        return self.forward_raw(x, keep=False)[0]

    def extra_repr(self) -> str:
        return f"input_dim={self.input_dim}, output_dim={self.output_dim}, num_layers={self.num_layers}, num_hiddens={self.num_hiddens}"


def _conv_out(n: int, pad: int, k: int, stride: int) -> int:
    """sbi ``calculate_filter_output_size`` (dilation 1)."""
    return (n + 2 * pad - (k - 1) - 1) // stride + 1


class CNNEmbedding(nn.Module):
    """``CNNEmbedding(input_shape, in_channels=1, out_channels_per_layer=[6, 12], num_conv_layers=2, num_linear_layers=2,
    num_linear_units=50, output_dim=20, kernel_size=5, pool_kernel_size=2)`` (sbi 0.22 signature, as called at
    /root/reference/modules.py:29-33 and floppityFUN.py:117-125): ``num_conv_layers x [Conv1d(stride 1, padding 1) -> ReLU
    -> MaxPool1d(pool_kernel_size)]`` on the spectrum as a single-channel signal, flattened into an ``FCEmbedding``."""

    def __init__(self, input_shape, in_channels: int = 1, out_channels_per_layer: Sequence[int] = (6, 12),
                 num_conv_layers: int = 2, num_linear_layers: int = 2, num_linear_units: int = 50, output_dim: int = 20,
                 kernel_size: int = 5, pool_kernel_size: int = 2, generator: Optional[torch.Generator] = None):
        super().__init__()
        input_shape = tuple(int(v) for v in (input_shape if isinstance(input_shape, (tuple, list)) else (input_shape,)))
        if len(input_shape) != 1:
            raise NotImplementedError("2-D inputs (images) are not on the FlopPITy path: spectra are 1-D")
        if int(in_channels) != 1:
            raise NotImplementedError("upstream reshapes the input to ONE channel (x.view(batch, 1, *input_shape))")
        if len(out_channels_per_layer) != int(num_conv_layers):
            raise AssertionError("out_channels_per_layer must have num_conv_layers entries")
        self.input_shape = input_shape
        self.input_dim = input_shape[0]
        self.channels = [1] + [int(c) for c in out_channels_per_layer]
        self.kernel_size, self.pool, self.pad = int(kernel_size), int(pool_kernel_size), 1
        self.num_conv_layers = int(num_conv_layers)
        self.lengths = [self.input_dim]                 # input length of each conv layer, then the flattened one
        self.conv_lengths = []
        for _ in range(self.num_conv_layers):
            lc = _conv_out(self.lengths[-1], self.pad, self.kernel_size, 1)
            lp = _conv_out(lc, 0, self.pool, self.pool)
            if lp < 1:
                raise ValueError("the input is too short for this many convolution / pooling layers")
            self.conv_lengths.append(lc)
            self.lengths.append(lp)
        for ci, co in zip(self.channels[:-1], self.channels[1:]):
            if ci * self.kernel_size > 64 or ci * co * self.kernel_size > 4096:
                raise NotImplementedError("convolution wider than the kernels cover (Cin*k <= 64, Cout*Cin*k <= 4096)")
        self.cw_off, self.cb_off, o = [], [], 0
        for ci, co in zip(self.channels[:-1], self.channels[1:]):
            self.cw_off.append(o); o += (co * ci * self.kernel_size + 3) // 4 * 4
            self.cb_off.append(o); o += (co + 3) // 4 * 4
        self.core = _FCCore(self.channels[-1] * self.lengths[-1], output_dim, num_linear_layers, num_linear_units, base=o)
        self.output_dim = self.core.output_dim
        self.param_count = o + self.core.count
        self._flat = nn.Parameter(torch.zeros(self.param_count, dtype=torch.float32))
        self.reset_parameters(generator)

    def conv_weight(self, i: int, flat: Optional[Tensor] = None) -> Tensor:
        flat = self._flat.detach() if flat is None else flat
        ci, co = self.channels[i], self.channels[i + 1]
        return flat[self.cw_off[i]: self.cw_off[i] + co * ci * self.kernel_size].view(co, ci, self.kernel_size)

    def conv_bias(self, i: int, flat: Optional[Tensor] = None) -> Tensor:
        flat = self._flat.detach() if flat is None else flat
        return flat[self.cb_off[i]: self.cb_off[i] + self.channels[i + 1]]

    @torch.no_grad()
    def reset_parameters(self, generator: Optional[torch.Generator] = None) -> None:
        """torch default Conv1d / Linear init: U(+-1/sqrt(fan_in)) for weights and biases."""
        self._flat.zero_()
        dev = self._flat.device
        for i in range(self.num_conv_layers):
            b = 1.0 / math.sqrt(self.channels[i] * self.kernel_size)
            w = self.conv_weight(i)
            w.copy_(((torch.rand(w.shape, generator=generator) * 2 - 1) * b).to(dev))
            self.conv_bias(i).copy_(((torch.rand(self.channels[i + 1], generator=generator) * 2 - 1) * b).to(dev))
        self.core.reset(self._flat, generator)

    def unpack(self, flat: Tensor, prefix: str = "") -> Dict[str, Tensor]:
        sd = {}
        for i in range(self.num_conv_layers):
            sd[f"{prefix}cnn_subnet.{3 * i}.weight"] = self.conv_weight(i, flat).clone()
            sd[f"{prefix}cnn_subnet.{3 * i}.bias"] = self.conv_bias(i, flat).clone()
        sd.update(self.core.unpack(flat, prefix + "linear_subnet."))
        return sd

    def state_dict(self, *args, destination=None, prefix: str = "", keep_vars: bool = False):
        sd = destination if destination is not None else {}
        sd.update(self.unpack(self._flat.detach(), prefix))
        return sd

    @torch.no_grad()
    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        for i in range(self.num_conv_layers):
            for name, dst in (("weight", self.conv_weight(i)), ("bias", self.conv_bias(i))):
                key = f"cnn_subnet.{3 * i}.{name}"
                if key in state_dict:
                    dst.copy_(state_dict[key].detach().to(dst.device, torch.float32))
                elif strict:
                    raise KeyError(f"missing key {key}")
        self.core.load(self._flat, state_dict, "linear_subnet.", strict)
        return torch.nn.modules.module._IncompatibleKeys([], [])

    def forward_raw(self, x: Tensor, keep: bool):
        """x [R, input_dim] (standardised spectra) -> (e [R, output_dim], saved activations when ``keep``)."""
        L, s = _lib.lib(), _lib.stream()
        if not x.is_cuda:
            raise RuntimeError("the embedding nets run on CUDA only; there is no CPU fallback")
        flat = self._flat.detach()
        R = x.shape[0]
        a = x.contiguous()
        ins, zs = [], []
        for i in range(self.num_conv_layers):
            ci, co = self.channels[i], self.channels[i + 1]
            z = torch.empty(R, co, self.conv_lengths[i], dtype=torch.float32, device=x.device)
            y = torch.empty(R, co, self.lengths[i + 1], dtype=torch.float32, device=x.device)
            _lib.check(L.fp_conv1d_relu_pool_fwd(_lib.ptr(a), _lib.ptr(self.conv_weight(i, flat).contiguous()),
                                                 _lib.ptr(self.conv_bias(i, flat)), _lib.ptr(z), _lib.ptr(y), R, ci, self.lengths[i],
                                                 co, self.kernel_size, self.pad, self.pool, s), "conv forward")
            ins.append(a); zs.append(z)
            a = y
        feat = a.reshape(R, -1)
        e, fc_saved = self.core.forward_raw(flat, feat, keep)
        return e, ((ins, zs, feat, fc_saved) if keep else None)

    def backward_raw(self, x: Tensor, saved, d_e: Tensor, grads: Tensor) -> None:
        L, s = _lib.lib(), _lib.stream()
        ins, zs, feat, fc_saved = saved
        flat = self._flat.detach()
        R = x.shape[0]
        d = self.core.backward_raw(flat, feat, fc_saved, d_e, grads, want_dx=True)
        for i in range(self.num_conv_layers - 1, -1, -1):
            ci, co = self.channels[i], self.channels[i + 1]
            dz = torch.empty_like(zs[i])
            dx = torch.empty_like(ins[i]) if i > 0 else None
            _lib.check(L.fp_conv1d_relu_pool_bwd(
                _lib.ptr(ins[i]), _lib.ptr(self.conv_weight(i, flat).contiguous()), _lib.ptr(zs[i]), _lib.ptr(d.contiguous()),
                _lib.ptr(dz), _lib.ptr(dx), _lib.ptr(grads[self.cw_off[i]:]), _lib.ptr(grads[self.cb_off[i]:]), R, ci,
                self.lengths[i], co, self.kernel_size, self.pad, self.pool, s), "conv backward")
            d = dx

    def forward(self, x: Tensor) -> Tensor:
        """Evaluation forward.  There is no CPU path: parameters that were never moved (the reference's multiNet keeps its
        sub-nets in a plain list, which ``.to(device)`` does not reach) are moved to the current CUDA device here."""
        dev = _lib.require_cuda(self._flat.device if self._flat.is_cuda else None)
        if not self._flat.is_cuda:
            self.to(dev)
        x = torch.as_tensor(x).to(dev, torch.float32).reshape(-1, self.input_dim)
        return self.forward_raw(x, keep=False)[0]

    def extra_repr(self) -> str:
        return (f"input_shape={self.input_shape}, channels={self.channels}, kernel_size={self.kernel_size}, "
                f"pool={self.pool}, linear={self.core.shapes}")


class FrozenEmbedding(nn.Module):
    """An embedding module WITHOUT trainable registered parameters, evaluated as it stands (no gradients): the reference's
    ``multiNet`` (floppityFUN.py:93-134) is one -- its CNNEmbedding sub-nets live in a plain list, so the optimiser upstream
    never sees them either."""

    def __init__(self, module: nn.Module, input_dim: int, output_dim: int):
        super().__init__()
        self.module = module
        self.input_dim, self.output_dim = int(input_dim), int(output_dim)
        self.param_count = 0
        self._flat = nn.Parameter(torch.zeros(0, dtype=torch.float32), requires_grad=False)

    def forward_raw(self, x: Tensor, keep: bool):
        with torch.no_grad():
            e = self.module(x)
        return e.to(x.device, torch.float32).contiguous(), None

    def backward_raw(self, x, saved, d_e, grads) -> None:
        return None

    def forward(self, x):
        return self.forward_raw(torch.as_tensor(x), False)[0]

    def state_dict(self, *args, destination=None, prefix: str = "", keep_vars: bool = False):
        return destination if destination is not None else {}

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        return torch.nn.modules.module._IncompatibleKeys([], [])


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


class PermutationInvariantEmbedding(nn.Module):
    """Imported by the reference (modules.py:1, floppityFUN.py:19) but never constructed there; kept so the import works."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError("PermutationInvariantEmbedding is imported but never used by FlopPITy; not implemented")
