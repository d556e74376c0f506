"""CPU oracle for the FlopPITy neural-posterior-estimation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``floppity_deprecated_b200/`` imports this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker / the CPU arm.

PARITY UNPINNED.  The arithmetic of this path does not live in /root/reference at
all: ``modules.py:63-64`` calls ``sbi.utils.get_nn_models.posterior_nn`` and
``sbi.inference.SNPE_C``; everything below those calls is sbi 0.22.0 / nflows 0.14
(versions inferred, the reference pins none -- README.md:13-20).  Neither package is
installed or installable here (no network), and the reference ships no tests, golden
vectors or fixtures.  This file is therefore a *restatement of the published
algorithms* (Durkan et al. 2019 neural spline flows as implemented by nflows; Greenberg
et al. 2019 APT/SNPE-C as implemented by sbi), op order kept so a later diff against a
real install is mechanical.  Call sites that anchor it in the reference:

    modules.py:57      utils.BoxUniform(low, high, device)            -> BoxUniform
    modules.py:63      posterior_nn(model='nsf', ...)                  -> posterior_nn / build_nsf
    modules.py:64      SNPE_C(prior, density_estimator, device)        -> SNPE_C
    floppity.py:216    inference.append_simulations(theta, x, proposal)
    floppity.py:219    inference.train(stop_after_epochs, num_atoms, force_first_round_loss, ...)
    floppity.py:225    inference.build_posterior(net).set_default_x(x)
    floppity.py:252    proposal.sample((n,))
    floppityFUN.py:35  posterior.log_prob(theta, x=default_x)

Everything is plain eager PyTorch on the CPU; ``.double()`` on any module gives the
fp64 ground truth used to classify fp32 bin flips.
"""
from __future__ import annotations

import copy
import math
from typing import Callable, Optional

import numpy as np
import torch
from torch import Tensor, nn
from torch.nn import functional as F

DEFAULT_MIN_BIN_WIDTH = 1e-3
DEFAULT_MIN_BIN_HEIGHT = 1e-3
DEFAULT_MIN_DERIVATIVE = 1e-3


# --------------------------------------------------------------------------------------
# nflows.utils.torchutils
# --------------------------------------------------------------------------------------
def searchsorted(bin_locations: Tensor, inputs: Tensor, eps: float = 1e-6) -> Tensor:
    """Compare-and-count bin search; the last knot is nudged by ``eps`` IN PLACE so an
    input sitting exactly on the right edge lands in bin K-1."""
    bin_locations[..., -1] += eps
    return torch.sum(inputs[..., None] >= bin_locations, dim=-1) - 1


def sum_except_batch(x: Tensor, num_batch_dims: int = 1) -> Tensor:
    reduce_dims = list(range(num_batch_dims, x.ndimension()))
    return torch.sum(x, dim=reduce_dims) if reduce_dims else x


def repeat_rows(x: Tensor, num_reps: int) -> Tensor:
    """[a, b] -> [a, a, ..., b, b, ...] (each row repeated ``num_reps`` times, consecutively)."""
    shape = x.shape
    x = x.unsqueeze(1)
    x = x.expand(shape[0], num_reps, *shape[1:])
    return x.reshape(shape[0] * num_reps, *shape[1:])


def create_alternating_binary_mask(features: int, even: bool = True) -> Tensor:
    mask = torch.zeros(features).byte()
    start = 0 if even else 1
    mask[start::2] += 1
    return mask


# --------------------------------------------------------------------------------------
# nflows.transforms.splines.rational_quadratic
# --------------------------------------------------------------------------------------
def rational_quadratic_spline(
    inputs, unnormalized_widths, unnormalized_heights, unnormalized_derivatives,
    inverse=False, left=0.0, right=1.0, bottom=0.0, top=1.0,
    min_bin_width=DEFAULT_MIN_BIN_WIDTH, min_bin_height=DEFAULT_MIN_BIN_HEIGHT,
    min_derivative=DEFAULT_MIN_DERIVATIVE, return_bins=False,
):
    if inputs.numel() and (torch.min(inputs) < left or torch.max(inputs) > right):
        raise ValueError("input outside spline domain")
    num_bins = unnormalized_widths.shape[-1]
    if min_bin_width * num_bins > 1.0:
        raise ValueError("Minimal bin width too large for the number of bins")
    if min_bin_height * num_bins > 1.0:
        raise ValueError("Minimal bin height too large for the number of bins")

    widths = F.softmax(unnormalized_widths, dim=-1)
    widths = min_bin_width + (1 - min_bin_width * num_bins) * widths
    cumwidths = torch.cumsum(widths, dim=-1)
    cumwidths = F.pad(cumwidths, pad=(1, 0), mode="constant", value=0.0)
    cumwidths = (right - left) * cumwidths + left
    cumwidths[..., 0] = left
    cumwidths[..., -1] = right
    widths = cumwidths[..., 1:] - cumwidths[..., :-1]

    derivatives = min_derivative + F.softplus(unnormalized_derivatives)

    heights = F.softmax(unnormalized_heights, dim=-1)
    heights = min_bin_height + (1 - min_bin_height * num_bins) * heights
    cumheights = torch.cumsum(heights, dim=-1)
    cumheights = F.pad(cumheights, pad=(1, 0), mode="constant", value=0.0)
    cumheights = (top - bottom) * cumheights + bottom
    cumheights[..., 0] = bottom
    cumheights[..., -1] = top
    heights = cumheights[..., 1:] - cumheights[..., :-1]

    if inverse:
        bin_idx = searchsorted(cumheights, inputs)[..., None]
    else:
        bin_idx = searchsorted(cumwidths, inputs)[..., None]

    input_cumwidths = cumwidths.gather(-1, bin_idx)[..., 0]
    input_bin_widths = widths.gather(-1, bin_idx)[..., 0]
    input_cumheights = cumheights.gather(-1, bin_idx)[..., 0]
    delta = heights / widths
    input_delta = delta.gather(-1, bin_idx)[..., 0]
    input_derivatives = derivatives.gather(-1, bin_idx)[..., 0]
    input_derivatives_plus_one = derivatives[..., 1:].gather(-1, bin_idx)[..., 0]
    input_heights = heights.gather(-1, bin_idx)[..., 0]

    if inverse:
        a = (inputs - input_cumheights) * (
            input_derivatives + input_derivatives_plus_one - 2 * input_delta
        ) + input_heights * (input_delta - input_derivatives)
        b = input_heights * input_derivatives - (inputs - input_cumheights) * (
            input_derivatives + input_derivatives_plus_one - 2 * input_delta
        )
        c = -input_delta * (inputs - input_cumheights)
        discriminant = b.pow(2) - 4 * a * c
        assert (discriminant >= 0).all()
        root = (2 * c) / (-b - torch.sqrt(discriminant))
        outputs = root * input_bin_widths + input_cumwidths
        theta_one_minus_theta = root * (1 - root)
        denominator = input_delta + (
            (input_derivatives + input_derivatives_plus_one - 2 * input_delta) * theta_one_minus_theta
        )
        derivative_numerator = input_delta.pow(2) * (
            input_derivatives_plus_one * root.pow(2)
            + 2 * input_delta * theta_one_minus_theta
            + input_derivatives * (1 - root).pow(2)
        )
        logabsdet = torch.log(derivative_numerator) - 2 * torch.log(denominator)
        out = (outputs, -logabsdet)
    else:
        theta = (inputs - input_cumwidths) / input_bin_widths
        theta_one_minus_theta = theta * (1 - theta)
        numerator = input_heights * (input_delta * theta.pow(2) + input_derivatives * theta_one_minus_theta)
        denominator = input_delta + (
            (input_derivatives + input_derivatives_plus_one - 2 * input_delta) * theta_one_minus_theta
        )
        outputs = input_cumheights + numerator / denominator
        derivative_numerator = input_delta.pow(2) * (
            input_derivatives_plus_one * theta.pow(2)
            + 2 * input_delta * theta_one_minus_theta
            + input_derivatives * (1 - theta).pow(2)
        )
        logabsdet = torch.log(derivative_numerator) - 2 * torch.log(denominator)
        out = (outputs, logabsdet)
    if return_bins:
        # also hand back the knot vector that was searched (for ulp-distance flip analysis)
        knots = cumheights if inverse else cumwidths
        return out + (bin_idx[..., 0], knots)
    return out


def unconstrained_rational_quadratic_spline(
    inputs, unnormalized_widths, unnormalized_heights, unnormalized_derivatives,
    inverse=False, tails="linear", tail_bound=1.0,
    min_bin_width=DEFAULT_MIN_BIN_WIDTH, min_bin_height=DEFAULT_MIN_BIN_HEIGHT,
    min_derivative=DEFAULT_MIN_DERIVATIVE, return_bins=False,
):
    inside_interval_mask = (inputs >= -tail_bound) & (inputs <= tail_bound)
    outside_interval_mask = ~inside_interval_mask

    outputs = torch.zeros_like(inputs)
    logabsdet = torch.zeros_like(inputs)
    bins = torch.full(inputs.shape, -1, dtype=torch.int64)
    knot_dist = torch.full(inputs.shape, float("inf"), dtype=inputs.dtype)

    if tails == "linear":
        unnormalized_derivatives = F.pad(unnormalized_derivatives, pad=(1, 1))
        constant = np.log(np.exp(1 - min_derivative) - 1)
        unnormalized_derivatives[..., 0] = constant
        unnormalized_derivatives[..., -1] = constant
        outputs[outside_interval_mask] = inputs[outside_interval_mask]
        logabsdet[outside_interval_mask] = 0
    else:
        raise RuntimeError("{} tails are not implemented.".format(tails))

    if torch.any(inside_interval_mask):
        res = rational_quadratic_spline(
            inputs=inputs[inside_interval_mask],
            unnormalized_widths=unnormalized_widths[inside_interval_mask, :],
            unnormalized_heights=unnormalized_heights[inside_interval_mask, :],
            unnormalized_derivatives=unnormalized_derivatives[inside_interval_mask, :],
            inverse=inverse, left=-tail_bound, right=tail_bound, bottom=-tail_bound, top=tail_bound,
            min_bin_width=min_bin_width, min_bin_height=min_bin_height, min_derivative=min_derivative,
            return_bins=return_bins,
        )
        outputs[inside_interval_mask], logabsdet[inside_interval_mask] = res[0], res[1]
        if return_bins:
            bins[inside_interval_mask] = res[2]
            x_in = inputs[inside_interval_mask]
            knot_dist[inside_interval_mask] = (x_in[..., None] - res[3]).abs().min(dim=-1).values
    if return_bins:
        return outputs, logabsdet, bins, knot_dist
    return outputs, logabsdet


# --------------------------------------------------------------------------------------
# nflows.nn.nets.resnet
# --------------------------------------------------------------------------------------
class ResidualBlock(nn.Module):
    def __init__(self, features, context_features, activation=F.relu, dropout_probability=0.0,
                 use_batch_norm=False, zero_initialization=True):
        super().__init__()
        if use_batch_norm:
            raise NotImplementedError("use_batch_norm=False on the FlopPITy path")
        self.activation = activation
        self.use_batch_norm = use_batch_norm
        if context_features is not None:
            self.context_layer = nn.Linear(context_features, features)
        self.linear_layers = nn.ModuleList([nn.Linear(features, features) for _ in range(2)])
        self.dropout = nn.Dropout(p=dropout_probability)
        if zero_initialization:
            nn.init.uniform_(self.linear_layers[-1].weight, -1e-3, 1e-3)
            nn.init.uniform_(self.linear_layers[-1].bias, -1e-3, 1e-3)

    def forward(self, inputs, context=None):
        temps = inputs
        temps = self.activation(temps)
        temps = self.linear_layers[0](temps)
        temps = self.activation(temps)
        temps = self.dropout(temps)
        temps = self.linear_layers[1](temps)
        if context is not None:
            temps = F.glu(torch.cat((temps, self.context_layer(context)), dim=1), dim=1)
        return inputs + temps


class ResidualNet(nn.Module):
    def __init__(self, in_features, out_features, hidden_features, context_features=None, num_blocks=2,
                 activation=F.relu, dropout_probability=0.0, use_batch_norm=False):
        super().__init__()
        self.hidden_features = hidden_features
        self.context_features = context_features
        if context_features is not None:
            self.initial_layer = nn.Linear(in_features + context_features, hidden_features)
        else:
            self.initial_layer = nn.Linear(in_features, hidden_features)
        self.blocks = nn.ModuleList([
            ResidualBlock(hidden_features, context_features, activation, dropout_probability, use_batch_norm)
            for _ in range(num_blocks)
        ])
        self.final_layer = nn.Linear(hidden_features, out_features)

    def forward(self, inputs, context=None):
        if context is None:
            temps = self.initial_layer(inputs)
        else:
            temps = self.initial_layer(torch.cat((inputs, context), dim=1))
        for block in self.blocks:
            temps = block(temps, context=context)
        return self.final_layer(temps)


# --------------------------------------------------------------------------------------
# nflows.transforms
# --------------------------------------------------------------------------------------
class PointwiseAffineTransform(nn.Module):
    def __init__(self, shift, scale):
        super().__init__()
        shift, scale = map(torch.as_tensor, (shift, scale))
        self.register_buffer("_shift", shift)
        self.register_buffer("_scale", scale)

    @property
    def _log_scale(self):
        return torch.log(torch.abs(self._scale))

    def _batch_logabsdet(self, batch_shape):
        if self._log_scale.numel() > 1:
            return self._log_scale.expand(batch_shape).sum()
        return self._log_scale * torch.Size(batch_shape).numel()

    def forward(self, inputs, context=None):
        num_batch, *batch_shape = inputs.size()
        outputs = inputs * self._scale + self._shift
        return outputs, self._batch_logabsdet(batch_shape).expand(num_batch)

    def inverse(self, inputs, context=None):
        num_batch, *batch_shape = inputs.size()
        outputs = (inputs - self._shift) / self._scale
        return outputs, -self._batch_logabsdet(batch_shape).expand(num_batch)


class PiecewiseRationalQuadraticCouplingTransform(nn.Module):
    """Coupling layer: features with mask>0 are transformed by an RQ spline whose
    parameters come from a ResidualNet of the other features (and the context)."""

    def __init__(self, mask, transform_net_create_fn, num_bins=10, tails="linear", tail_bound=1.0,
                 min_bin_width=DEFAULT_MIN_BIN_WIDTH, min_bin_height=DEFAULT_MIN_BIN_HEIGHT,
                 min_derivative=DEFAULT_MIN_DERIVATIVE):
        super().__init__()
        mask = torch.as_tensor(mask)
        self.features = len(mask)
        fv = torch.arange(self.features)
        self.register_buffer("identity_features", fv.masked_select(mask <= 0))
        self.register_buffer("transform_features", fv.masked_select(mask > 0))
        self.num_bins, self.tails, self.tail_bound = num_bins, tails, tail_bound
        self.min_bin_width, self.min_bin_height, self.min_derivative = min_bin_width, min_bin_height, min_derivative
        mult = num_bins * 3 - 1 if tails == "linear" else num_bins * 3 + 1
        self.transform_net = transform_net_create_fn(len(self.identity_features), len(self.transform_features) * mult)
        self.last_bins = None  # filled when record_bins is on (test hook only)
        self.record_bins = False

    def _piecewise_cdf(self, inputs, transform_params, inverse=False):
        K = self.num_bins
        unnormalized_widths = transform_params[..., :K]
        unnormalized_heights = transform_params[..., K:2 * K]
        unnormalized_derivatives = transform_params[..., 2 * K:]
        if hasattr(self.transform_net, "hidden_features"):
            unnormalized_widths = unnormalized_widths / np.sqrt(self.transform_net.hidden_features)
            unnormalized_heights = unnormalized_heights / np.sqrt(self.transform_net.hidden_features)
        res = unconstrained_rational_quadratic_spline(
            inputs=inputs, unnormalized_widths=unnormalized_widths, unnormalized_heights=unnormalized_heights,
            unnormalized_derivatives=unnormalized_derivatives, inverse=inverse, tails=self.tails,
            tail_bound=self.tail_bound, min_bin_width=self.min_bin_width, min_bin_height=self.min_bin_height,
            min_derivative=self.min_derivative, return_bins=self.record_bins,
        )
        if self.record_bins:
            self.last_bins = (res[2], res[3])
        return res[0], res[1]

    def _coupling(self, inputs, context, inverse):
        identity_split = inputs[:, self.identity_features]
        transform_split = inputs[:, self.transform_features]
        transform_params = self.transform_net(identity_split, context)
        b, d = transform_split.shape
        transform_params = transform_params.reshape(b, d, -1)
        transform_split, logabsdet = self._piecewise_cdf(transform_split, transform_params, inverse)
        logabsdet = sum_except_batch(logabsdet)
        outputs = torch.empty_like(inputs)
        outputs[:, self.identity_features] = identity_split
        outputs[:, self.transform_features] = transform_split
        return outputs, logabsdet

    def forward(self, inputs, context=None):
        return self._coupling(inputs, context, inverse=False)

    def inverse(self, inputs, context=None):
        return self._coupling(inputs, context, inverse=True)


class LULinear(nn.Module):
    def __init__(self, features, identity_init=True, eps=1e-3):
        super().__init__()
        self.features = features
        self.eps = eps
        self.bias = nn.Parameter(torch.zeros(features))
        self.lower_indices = np.tril_indices(features, k=-1)
        self.upper_indices = np.triu_indices(features, k=1)
        self.diag_indices = np.diag_indices(features)
        n_tri = ((features - 1) * features) // 2
        self.lower_entries = nn.Parameter(torch.zeros(n_tri))
        self.upper_entries = nn.Parameter(torch.zeros(n_tri))
        self.unconstrained_upper_diag = nn.Parameter(torch.zeros(features))
        nn.init.zeros_(self.bias)
        if identity_init:
            nn.init.zeros_(self.lower_entries)
            nn.init.zeros_(self.upper_entries)
            nn.init.constant_(self.unconstrained_upper_diag, np.log(np.exp(1 - self.eps) - 1))
        else:
            stdv = 1.0 / np.sqrt(features)
            nn.init.uniform_(self.lower_entries, -stdv, stdv)
            nn.init.uniform_(self.upper_entries, -stdv, stdv)
            nn.init.uniform_(self.unconstrained_upper_diag, -stdv, stdv)

    @property
    def upper_diag(self):
        return F.softplus(self.unconstrained_upper_diag) + self.eps

    def _create_lower_upper(self):
        lower = self.lower_entries.new_zeros(self.features, self.features)
        lower[self.lower_indices[0], self.lower_indices[1]] = self.lower_entries
        lower[self.diag_indices[0], self.diag_indices[1]] = 1.0
        upper = self.upper_entries.new_zeros(self.features, self.features)
        upper[self.upper_indices[0], self.upper_indices[1]] = self.upper_entries
        upper[self.diag_indices[0], self.diag_indices[1]] = self.upper_diag
        return lower, upper

    def logabsdet(self):
        return torch.sum(torch.log(self.upper_diag))

    def forward(self, inputs, context=None):
        lower, upper = self._create_lower_upper()
        outputs = F.linear(inputs, upper)
        outputs = F.linear(outputs, lower, self.bias)
        return outputs, self.logabsdet() * inputs.new_ones(outputs.shape[0])

    def inverse(self, inputs, context=None):
        lower, upper = self._create_lower_upper()
        outputs = inputs - self.bias
        outputs = torch.linalg.solve_triangular(lower, outputs.t(), upper=False, unitriangular=True)
        outputs = torch.linalg.solve_triangular(upper, outputs, upper=True, unitriangular=False)
        outputs = outputs.t()
        return outputs, -self.logabsdet() * inputs.new_ones(outputs.shape[0])


class CompositeTransform(nn.Module):
    def __init__(self, transforms):
        super().__init__()
        self._transforms = nn.ModuleList(transforms)

    @staticmethod
    def _cascade(inputs, funcs, context):
        outputs = inputs
        total_logabsdet = inputs.new_zeros(inputs.shape[0])
        for func in funcs:
            outputs, logabsdet = func(outputs, context)
            total_logabsdet = total_logabsdet + logabsdet
        return outputs, total_logabsdet

    def forward(self, inputs, context=None):
        return self._cascade(inputs, list(self._transforms), context)

    def inverse(self, inputs, context=None):
        return self._cascade(inputs, [t.inverse for t in self._transforms[::-1]], context)


class StandardNormal(nn.Module):
    def __init__(self, shape):
        super().__init__()
        self._shape = torch.Size(shape)
        self.register_buffer(
            "_log_z", torch.tensor(0.5 * np.prod(shape) * np.log(2 * np.pi), dtype=torch.float64), persistent=False)

    def log_prob(self, inputs, context=None):
        neg_energy = -0.5 * sum_except_batch(inputs ** 2, num_batch_dims=1)
        return neg_energy - self._log_z

    def sample(self, num_samples, context=None, noise: Optional[Tensor] = None):
        if noise is not None:
            return noise
        return torch.randn(num_samples, *self._shape, device=self._log_z.device)


class Standardize(nn.Module):
    def __init__(self, mean, std):
        super().__init__()
        mean, std = map(torch.as_tensor, (mean, std))
        self.register_buffer("_mean", mean)
        self.register_buffer("_std", std)

    def forward(self, tensor):
        return (tensor - self._mean) / self._std


class Flow(nn.Module):
    def __init__(self, transform, distribution, embedding_net=None):
        super().__init__()
        self._transform = transform
        self._distribution = distribution
        self._embedding_net = embedding_net if embedding_net is not None else nn.Identity()

    def log_prob(self, inputs, context=None):
        inputs = torch.as_tensor(inputs)
        if context is not None:
            context = torch.as_tensor(context)
            if inputs.shape[0] != context.shape[0]:
                raise ValueError("Number of input items must be equal to number of context items.")
        embedded_context = self._embedding_net(context)
        noise, logabsdet = self._transform(inputs, context=embedded_context)
        log_prob = self._distribution.log_prob(noise, context=embedded_context)
        return log_prob + logabsdet

    def sample(self, num_samples, context=None, noise: Optional[Tensor] = None):
        """``noise`` ([M*num_samples, D]) replaces the base draw so parity never depends on RNG streams."""
        embedded_context = self._embedding_net(context)
        M = embedded_context.shape[0]
        z = self._distribution.sample(M * num_samples, noise=noise)
        embedded_context = repeat_rows(embedded_context, num_reps=num_samples)
        samples, _ = self._transform.inverse(z, context=embedded_context)
        return samples.reshape(M, num_samples, -1)


# --------------------------------------------------------------------------------------
# sbi.utils.sbiutils / sbi.neural_nets.flow
# --------------------------------------------------------------------------------------
def _valid_rows(t: Tensor) -> Tensor:
    flat = t.reshape(t.shape[0], -1)
    return ~(torch.isnan(flat).any(dim=1) | torch.isinf(flat).any(dim=1))


def standardizing_transform(batch_t: Tensor, min_std: float = 1e-14) -> PointwiseAffineTransform:
    ok = _valid_rows(batch_t)
    t_mean = torch.mean(batch_t[ok], dim=0)
    t_std = torch.std(batch_t[ok], dim=0)
    t_std[t_std < min_std] = min_std
    return PointwiseAffineTransform(shift=-t_mean / t_std, scale=1 / t_std)


def standardizing_net(batch_t: Tensor, min_std: float = 1e-7) -> Standardize:
    ok = _valid_rows(batch_t)
    t_mean = torch.mean(batch_t[ok], dim=0)
    t_std = torch.std(batch_t[ok], dim=0)
    t_std[t_std < min_std] = min_std
    return Standardize(t_mean, t_std)


class FCEmbedding(nn.Module):
    """[UPSTREAM-RECALL sbi 0.22 ``sbi.neural_nets.embedding_nets.FCEmbedding``] as FlopPITy builds it at
    /root/reference/modules.py:25 (``FCEmbedding(input_dim=nwvl, num_hiddens=256, num_layers=4, output_dim=embed_size)``):
    Linear(input, hidden)+ReLU, (num_layers-2) x [Linear(hidden, hidden)+ReLU], Linear(hidden, output)+ReLU."""

    def __init__(self, input_dim: int, output_dim: int = 20, num_layers: int = 2, num_hiddens: int = 50):
        super().__init__()
        layers = [nn.Linear(input_dim, num_hiddens), nn.ReLU()]
        for _ in range(num_layers - 2):
            layers.append(nn.Linear(num_hiddens, num_hiddens))
            layers.append(nn.ReLU())
        layers.append(nn.Linear(num_hiddens, output_dim))
        layers.append(nn.ReLU())
        self.net = nn.Sequential(*layers)

    def forward(self, x: Tensor) -> Tensor:
        return self.net(x)


def calculate_filter_output_size(input_size, padding, dilation, kernel, stride) -> int:
    """[UPSTREAM-RECALL sbi 0.22 embedding_nets.cnn]"""
    return int((int(input_size) + 2 * int(padding) - int(dilation) * (int(kernel) - 1) - 1) / int(stride) + 1)


class CNNEmbedding(nn.Module):
    """[UPSTREAM-RECALL sbi 0.22 ``sbi.neural_nets.embedding_nets.CNNEmbedding``, 1-D case] as FlopPITy builds it at
    /root/reference/modules.py:29-33 (-embedding_type CNN) and floppityFUN.py:117-125 (multiNet): num_conv_layers x
    [Conv1d(stride 1, padding 1), ReLU, MaxPool1d(pool_kernel_size)] on the input viewed as ONE channel, flattened into
    ``FCEmbedding(out_channels[-1] * length, output_dim, num_linear_layers, num_linear_units)``."""

    def __init__(self, input_shape, in_channels=1, out_channels_per_layer=(6, 12), num_conv_layers=2, num_linear_layers=2,
                 num_linear_units=50, output_dim=20, kernel_size=5, pool_kernel_size=2):
        super().__init__()
        assert len(input_shape) == 1, "1-D spectra on the FlopPITy path"
        assert len(out_channels_per_layer) == num_conv_layers
        self.input_shape = tuple(input_shape)
        stride, padding = 1, 1
        layers, size = [], int(input_shape[0])
        for ii in range(num_conv_layers):
            layers += [nn.Conv1d(in_channels if ii == 0 else out_channels_per_layer[ii - 1], out_channels_per_layer[ii],
                                 kernel_size=kernel_size, stride=stride, padding=padding),
                       nn.ReLU(inplace=True), nn.MaxPool1d(kernel_size=pool_kernel_size)]
            size = calculate_filter_output_size(size, padding, 1, kernel_size, stride)
            size = calculate_filter_output_size(size, 0, 1, pool_kernel_size, pool_kernel_size)
        self.cnn_subnet = nn.Sequential(*layers)
        self.linear_subnet = FCEmbedding(input_dim=out_channels_per_layer[-1] * size, output_dim=output_dim,
                                         num_layers=num_linear_layers, num_hiddens=num_linear_units)

    def forward(self, x: Tensor) -> Tensor:
        batch_size = x.size(0)
        x = self.cnn_subnet(x.view(batch_size, 1, *self.input_shape))
        return self.linear_subnet(x.view(batch_size, -1))


class MultiNet(nn.Module):
    """Restatement of the reference's own ``multiNet`` (/root/reference/floppityFUN.py:93-134): one CNNEmbedding per
    observation on consecutive column slices, outputs concatenated.  As in the reference the sub-nets sit in a plain list:
    their parameters are NOT registered (never trained, not part of the state_dict)."""

    def __init__(self, nwvl, make_net):
        super().__init__()
        self.nwvl = [int(n) for n in nwvl]
        self.nets = [make_net(i, n) for i, n in enumerate(self.nwvl)]

    def forward(self, X: Tensor) -> Tensor:
        out, o = [], 0
        for n, net in zip(self.nwvl, self.nets):
            out.append(net(X[:, o:o + n]))
            o += n
        return torch.cat(out, dim=1)


def build_nsf(batch_x: Tensor, batch_y: Tensor, z_score_x="independent", z_score_y="independent",
              hidden_features=50, num_transforms=5, num_bins=10, embedding_net: nn.Module = None,
              tail_bound=3.0, hidden_layers_spline_context=1, num_blocks=2, dropout_probability=0.0,
              use_batch_norm=False, **kwargs) -> Flow:
    """sbi naming: ``batch_x`` is theta, ``batch_y`` is the spectrum (modules.py:63 via posterior_nn)."""
    embedding_net = embedding_net if embedding_net is not None else nn.Identity()
    x_numel = batch_x[0].numel()
    y_numel = embedding_net(batch_y[:1]).numel()
    if x_numel == 1:
        raise NotImplementedError("1-D theta uses a different conditioner upstream; not on the FlopPITy path")

    def conditioner(in_features, out_features):
        return ResidualNet(in_features, out_features, hidden_features=hidden_features, context_features=y_numel,
                           num_blocks=num_blocks, activation=F.relu, dropout_probability=dropout_probability,
                           use_batch_norm=use_batch_norm)

    transform_list = []
    for i in range(num_transforms):
        transform_list.append(PiecewiseRationalQuadraticCouplingTransform(
            mask=create_alternating_binary_mask(x_numel, even=(i % 2 == 0)),
            transform_net_create_fn=conditioner, num_bins=num_bins, tails="linear", tail_bound=tail_bound))
        transform_list.append(LULinear(x_numel, identity_init=True))
    if z_score_x not in (None, "none", False):
        transform_list = [standardizing_transform(batch_x)] + transform_list
    if z_score_y not in (None, "none", False):
        embedding_net = nn.Sequential(standardizing_net(batch_y), embedding_net)
    return Flow(CompositeTransform(transform_list), StandardNormal((x_numel,)), embedding_net)


def posterior_nn(model: str, z_score_theta="independent", z_score_x="independent", hidden_features=50,
                 num_transforms=5, num_bins=10, embedding_net: nn.Module = None, **kwargs) -> Callable:
    if model != "nsf":
        raise NotImplementedError("only model='nsf' is on the FlopPITy hot path (floppity.py:43)")

    def build_fn(batch_theta, batch_x):
        return build_nsf(batch_x=batch_theta, batch_y=batch_x, z_score_x=z_score_theta, z_score_y=z_score_x,
                         hidden_features=hidden_features, num_transforms=num_transforms, num_bins=num_bins,
                         embedding_net=embedding_net, **kwargs)

    return build_fn


# --------------------------------------------------------------------------------------
# sbi.utils.BoxUniform and support checks
# --------------------------------------------------------------------------------------
class BoxUniform:
    def __init__(self, low, high, device="cpu"):
        self.low = torch.as_tensor(low, dtype=torch.float32).reshape(-1)
        self.high = torch.as_tensor(high, dtype=torch.float32).reshape(-1)

    def within(self, theta: Tensor) -> Tensor:
        return ((theta >= self.low) & (theta <= self.high)).all(dim=-1)

    def log_prob(self, theta: Tensor) -> Tensor:
        lp = -torch.log(self.high - self.low).sum()
        inside = ((theta >= self.low) & (theta < self.high)).all(dim=-1)  # torch Uniform: low <= v < high
        return torch.where(inside, lp.expand(theta.shape[:-1]), torch.full(theta.shape[:-1], -float("inf")))

    def sample(self, sample_shape=torch.Size()):
        u = torch.rand(*sample_shape, self.low.numel())
        return self.low + u * (self.high - self.low)


# --------------------------------------------------------------------------------------
# sbi.inference.snpe.snpe_c  (atomic APT loss) and the PosteriorEstimator.train loop
# --------------------------------------------------------------------------------------
def log_prob_proposal_posterior_atomic(net: Flow, prior: BoxUniform, theta: Tensor, x: Tensor, masks: Tensor,
                                       num_atoms: int, use_combined_loss: bool,
                                       choices: Optional[Tensor] = None) -> Tensor:
    """floppity.py:219-221 passes num_atoms / use_combined_loss.  ``choices`` ([B, A-1] int64, no
    self-index) may be injected so the contrastive draw is shared with the implementation under test."""
    batch_size = theta.shape[0]
    num_atoms = int(min(max(num_atoms, 2), batch_size))
    repeated_x = repeat_rows(x, num_atoms)
    if choices is None:
        probs = torch.ones(batch_size, batch_size) * (1 - torch.eye(batch_size)) / (batch_size - 1)
        choices = torch.multinomial(probs, num_samples=num_atoms - 1, replacement=False)
    contrasting_theta = theta[choices]
    atomic_theta = torch.cat((theta[:, None, :], contrasting_theta), dim=1).reshape(batch_size * num_atoms, -1)
    log_prob_posterior = net.log_prob(atomic_theta, repeated_x).reshape(batch_size, num_atoms)
    assert torch.isfinite(log_prob_posterior).all()
    log_prob_prior = prior.log_prob(atomic_theta).reshape(batch_size, num_atoms)
    assert torch.isfinite(log_prob_prior).all()
    unnormalized_log_prob = log_prob_posterior - log_prob_prior
    log_prob_proposal_posterior = unnormalized_log_prob[:, 0] - torch.logsumexp(unnormalized_log_prob, dim=-1)
    if use_combined_loss:
        log_prob_posterior_non_atomic = net.log_prob(theta, x)
        log_prob_proposal_posterior = masks.reshape(-1) * log_prob_posterior_non_atomic + log_prob_proposal_posterior
    return log_prob_proposal_posterior


class SNPE_C:
    """Trainer loop of sbi 0.22 ``PosteriorEstimator.train`` as invoked at floppity.py:216-221."""

    def __init__(self, prior: BoxUniform, density_estimator: Callable, device: str = "cpu"):
        self._prior = prior
        self._build_neural_net = density_estimator
        self._neural_net = None
        self._theta_roundwise, self._x_roundwise, self._prior_masks, self._data_round_index = [], [], [], []
        self._proposal_roundwise = []
        self._round = 0
        self._val_log_prob = float("-inf")
        self._summary = dict(epochs_trained=[], best_validation_log_prob=[], validation_log_probs=[],
                             training_log_probs=[], epoch_durations_sec=[])

    def append_simulations(self, theta: Tensor, x: Tensor, proposal=None):
        theta = torch.as_tensor(theta, dtype=torch.float32)
        x = torch.as_tensor(x, dtype=torch.float32)
        if proposal is None or proposal is self._prior:
            current_round = 0
        else:
            current_round = max(self._data_round_index) + 1 if self._data_round_index else 1
        if current_round == 0:
            ok = _valid_rows(x)
            theta, x = theta[ok], x[ok]
        prior_masks = torch.ones(theta.shape[0], 1) if current_round == 0 else torch.zeros(theta.shape[0], 1)
        self._theta_roundwise.append(theta)
        self._x_roundwise.append(x)
        self._prior_masks.append(prior_masks)
        self._proposal_roundwise.append(proposal)
        self._data_round_index.append(current_round)
        return self

    def _converged(self, epoch: int, stop_after_epochs: int) -> bool:
        converged = False
        if epoch == 0 or self._val_log_prob > self._best_val_log_prob:
            self._best_val_log_prob = self._val_log_prob
            self._epochs_since_last_improvement = 0
            self._best_model_state_dict = copy.deepcopy(self._neural_net.state_dict())
        else:
            self._epochs_since_last_improvement += 1
        if self._epochs_since_last_improvement > stop_after_epochs - 1:
            self._neural_net.load_state_dict(self._best_model_state_dict)
            converged = True
        return converged

    def _loss(self, theta, x, masks, force_first_round_loss, choices=None):
        if self._round == 0 or force_first_round_loss:
            log_prob = self._neural_net.log_prob(theta, x)
        else:
            log_prob = log_prob_proposal_posterior_atomic(self._neural_net, self._prior, theta, x, masks,
                                                          self._num_atoms, self._use_combined_loss, choices)
        return -log_prob

    def train(self, num_atoms=10, training_batch_size=50, learning_rate=5e-4, validation_fraction=0.1,
              stop_after_epochs=20, max_num_epochs=2 ** 31 - 1, clip_max_norm=5.0, force_first_round_loss=False,
              discard_prior_samples=False, use_combined_loss=False, retrain_from_scratch=False,
              show_train_summary=False, generator: Optional[torch.Generator] = None):
        self._num_atoms, self._use_combined_loss = num_atoms, use_combined_loss
        self._round = max(self._data_round_index)
        start_idx = int(discard_prior_samples and self._round > 0)
        theta = torch.cat(self._theta_roundwise[start_idx:])
        x = torch.cat(self._x_roundwise[start_idx:])
        masks = torch.cat(self._prior_masks[start_idx:])
        n = theta.shape[0]
        n_train = int((1 - validation_fraction) * n)
        perm = torch.randperm(n, generator=generator)
        train_idx, val_idx = perm[:n_train], perm[n_train:]
        bs_train = min(training_batch_size, n_train)
        bs_val = min(training_batch_size, n - n_train)

        if self._neural_net is None or retrain_from_scratch:
            self._neural_net = self._build_neural_net(theta[train_idx], x[train_idx])
        optimizer = torch.optim.Adam(list(self._neural_net.parameters()), lr=learning_rate)
        epoch, self._val_log_prob = 0, float("-inf")

        while epoch <= max_num_epochs and not self._converged(epoch, stop_after_epochs):
            self._neural_net.train()
            order = train_idx[torch.randperm(n_train, generator=generator)]
            train_sum, nb = 0.0, n_train // bs_train  # drop_last=True
            for b in range(nb):
                idx = order[b * bs_train:(b + 1) * bs_train]
                optimizer.zero_grad()
                losses = self._loss(theta[idx], x[idx], masks[idx], force_first_round_loss)
                loss = torch.mean(losses)
                train_sum += losses.sum().item()
                loss.backward()
                if clip_max_norm is not None:
                    torch.nn.utils.clip_grad_norm_(self._neural_net.parameters(), max_norm=clip_max_norm)
                optimizer.step()
            epoch += 1
            self._summary["training_log_probs"].append(-train_sum / max(nb * bs_train, 1))
            self._neural_net.eval()
            val_sum, nvb = 0.0, (n - n_train) // bs_val
            with torch.no_grad():
                order_v = val_idx[torch.randperm(n - n_train, generator=generator)]
                for b in range(nvb):
                    idx = order_v[b * bs_val:(b + 1) * bs_val]
                    val_sum += self._loss(theta[idx], x[idx], masks[idx], force_first_round_loss).sum().item()
            self._val_log_prob = -val_sum / max(nvb * bs_val, 1)
            self._summary["validation_log_probs"].append(self._val_log_prob)
        self._summary["epochs_trained"].append(epoch)
        self._summary["best_validation_log_prob"].append(self._best_val_log_prob)
        self._neural_net.zero_grad(set_to_none=True)
        return copy.deepcopy(self._neural_net)

    def build_posterior(self, density_estimator: Optional[Flow] = None):
        net = density_estimator if density_estimator is not None else self._neural_net
        return DirectPosterior(net, self._prior)


# --------------------------------------------------------------------------------------
# sbi.inference.posteriors.direct_posterior + sbi.samplers.rejection
# --------------------------------------------------------------------------------------
class DirectPosterior:
    def __init__(self, posterior_estimator: Flow, prior: BoxUniform):
        self.posterior_estimator = posterior_estimator
        self.prior = prior
        self._x = None
        self._leakage_cache = None

    def set_default_x(self, x):
        self._x = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x, dtype=torch.float32).reshape(1, -1)
        self._leakage_cache = None
        return self

    def _x_or_default(self, x):
        if x is None:
            if self._x is None:
                raise ValueError("no x given and no default x set")
            return self._x
        return torch.as_tensor(x, dtype=torch.float32).reshape(1, -1)

    def sample(self, sample_shape=torch.Size(), x=None, max_sampling_batch_size: int = 10_000) -> Tensor:
        num_samples = int(torch.Size(sample_shape).numel())
        x = self._x_or_default(x)
        samples, _ = self._accept_reject(num_samples, x, max_sampling_batch_size)
        return samples.reshape(*sample_shape, -1)

    @torch.no_grad()
    def _accept_reject(self, num_samples, x, max_sampling_batch_size=10_000):
        self.posterior_estimator.eval()
        accepted, num_remaining, num_total = [], num_samples, 0
        bs = min(num_samples, max_sampling_batch_size)
        acceptance_rate = 1.0
        while num_remaining > 0:
            cand = self.posterior_estimator.sample(bs, context=x).reshape(bs, -1)
            ok = self.prior.within(cand)
            accepted.append(cand[ok])
            num_remaining -= int(ok.sum())
            num_total += bs
            acceptance_rate = (num_samples - num_remaining) / num_total
            bs = min(max_sampling_batch_size, max(int(1.5 * num_remaining / max(acceptance_rate, 1e-12)), 100))
        return torch.cat(accepted)[:num_samples], torch.as_tensor(acceptance_rate)

    @torch.no_grad()
    def leakage_correction(self, x, num_rejection_samples: int = 10_000) -> Tensor:
        if self._leakage_cache is None:
            _, acc = self._accept_reject(num_rejection_samples, x)
            self._leakage_cache = acc
        return self._leakage_cache

    def log_prob(self, theta, x=None, norm_posterior: bool = True) -> Tensor:
        theta = torch.as_tensor(np.asarray(theta) if not torch.is_tensor(theta) else theta, dtype=torch.float32)
        theta = theta.reshape(-1, theta.shape[-1])
        x = self._x_or_default(x)
        self.posterior_estimator.eval()
        with torch.no_grad():
            lp = self.posterior_estimator.log_prob(theta, x.expand(theta.shape[0], -1).contiguous())
            lp = torch.where(self.prior.within(theta), lp, torch.full_like(lp, -float("inf")))
            if norm_posterior:
                lp = lp - torch.log(self.leakage_correction(x))
        return lp


# --------------------------------------------------------------------------------------
# helpers used by tests / bench (not upstream API)
# --------------------------------------------------------------------------------------
def coupling_layers(flow: Flow):
    return [t for t in flow._transform._transforms if isinstance(t, PiecewiseRationalQuadraticCouplingTransform)]


def set_record_bins(flow: Flow, on: bool = True):
    for c in coupling_layers(flow):
        c.record_bins = on
        c.last_bins = None


def randomize_lu(flow: Flow, generator: torch.Generator, scale: float = 0.2):
    """LULinear is the identity at init; tests perturb it so the LU path is actually exercised."""
    with torch.no_grad():
        for t in flow._transform._transforms:
            if isinstance(t, LULinear):
                for p in (t.lower_entries, t.upper_entries, t.unconstrained_upper_diag, t.bias):
                    p.add_(scale * torch.randn(p.shape, generator=generator, dtype=p.dtype))


class InjectedDropout(nn.Module):
    """``nn.Dropout`` with the keep mask handed in (inverted dropout: kept activations are scaled by 1/(1-p)), so a
    parity test can give the oracle the very mask the CUDA kernels drew (their RNG is not torch's).  Identity in
    ``eval()`` mode, like ``nn.Dropout``."""

    def __init__(self, keep: Tensor, p: float):
        super().__init__()
        self.keep, self.p = keep, float(p)

    def forward(self, x):
        if not self.training:
            return x
        return x * self.keep.to(x.dtype) / (1.0 - self.p)


def inject_dropout_masks(flow: Flow, keep_masks, p: float):
    """``keep_masks[layer][block]`` ([rows, hidden] 0/1) replaces the dropout module of every residual block."""
    for i, c in enumerate(coupling_layers(flow)):
        for k, blk in enumerate(c.transform_net.blocks):
            blk.dropout = InjectedDropout(keep_masks[i][k], p)
