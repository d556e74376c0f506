"""sbi >= 0.23 spelling of the same estimator (``DensityEstimator.log_prob / sample / loss``), named in
BASELINE.json ``north_star``; the reference itself uses the 0.22 ``Flow`` spelling (modules.py:63).

Shapes follow sbi's ``NFlowsFlow``: ``input`` is ``(sample_dim, batch_dim, D)``, ``condition`` is
``(batch_dim, C)``.  The condition is broadcast over ``sample_dim`` WITHOUT materialising the repeat: the
kernels take one context row per ``sample_dim`` group (the context-hoisting of SURVEY.md rows A8/A10).
"""
from __future__ import annotations

import torch
from torch import Tensor, nn

from .flow import NSFFlow


class NSFDensityEstimator(nn.Module):
    def __init__(self, net: NSFFlow):
        super().__init__()
        self.net = net
        self.input_shape = torch.Size([net.D])
        self.condition_shape = torch.Size([net.C])

    @property
    def embedding_net(self):
        return nn.Identity()

    def log_prob(self, input: Tensor, condition: Tensor) -> Tensor:
        """input (sample_dim, batch_dim, D), condition (batch_dim, C) -> (sample_dim, batch_dim)."""
        if input.dim() == 2:
            input = input.unsqueeze(0)
        S, B, D = input.shape
        if condition.shape[0] != B:
            raise ValueError("batch_dim of input and condition differ")
        # rows ordered (batch, sample) so that each context row serves S consecutive flow rows
        rows = input.permute(1, 0, 2).reshape(B * S, D)
        lp = self.net.log_prob(rows, condition)          # Rc = B, R = B*S -> context row r // S
        return lp.reshape(B, S).permute(1, 0)

    def loss(self, input: Tensor, condition: Tensor) -> Tensor:
        """input (batch_dim, D) -> (batch_dim,) negative log-prob."""
        return -self.net.log_prob(input, condition)

    def sample(self, sample_shape, condition: Tensor) -> Tensor:
        """-> (*sample_shape, batch_dim, D)."""
        n = int(torch.Size(sample_shape).numel())
        out = self.net.sample(n, context=condition)        # (batch, n, D)
        return out.permute(1, 0, 2).reshape(*torch.Size(sample_shape), condition.shape[0], self.net.D)

    def sample_and_log_prob(self, sample_shape, condition: Tensor):
        s = self.sample(sample_shape, condition)
        flat = s.reshape(-1, condition.shape[0], self.net.D)
        return s, self.log_prob(flat, condition).reshape(*torch.Size(sample_shape), condition.shape[0])
