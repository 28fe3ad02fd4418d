"""Oracle (test infrastructure): the reference's LoRA-injected projection.

Restates ``modules/lora.py:12-27`` (``get_lora``) and the arithmetic of the
third-party package it wraps, ``loralib==0.1`` (``requirements.txt:16``), which
is not vendored in ``/root/reference`` and not installed here.  Published
loralib-0.1 semantics restated below:

* ``loralib.Linear(in, out, r, lora_alpha, lora_dropout)``:
  ``lora_A [r,in]`` kaiming-uniform(a=sqrt 5), ``lora_B [out,r]`` zeros,
  ``scaling = lora_alpha / r``; train-mode forward
  ``F.linear(x, W, b) + (dropout(x) @ A.T @ B.T) * scaling``.
* ``loralib.Conv2d(in, out, k, r, lora_alpha, lora_dropout)``:
  ``lora_A [r*k, in*k]``, ``lora_B [out*k, r*k]``; forward
  ``conv2d(x, W + (B @ A).view(W.shape) * scaling, b)`` with the default
  stride/padding (``modules/lora.py:16`` forwards only ``kernel_size[0]``).

PARITY UNPINNED by the reference (no tests / golden vectors, package absent);
cross-checked by closed-form identities in ``tests/test_oracle.py``.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn


class RefLoRALinear(nn.Module):
    """loralib-0.1 ``Linear`` as adapted by ``modules/lora.py:12-27``."""

    def __init__(self, in_features, out_features, rank=4, alpha=1, dropout=0.0):
        super().__init__()
        self.in_features, self.out_features, self.r = in_features, out_features, rank
        # nn.Linear.__init__ creates weight/bias; get_lora overwrites them (lora.py:20-21)
        self.weight = nn.Parameter(torch.empty(out_features, in_features), requires_grad=False)
        self.bias = None
        self.lora_A = nn.Parameter(torch.zeros(rank, in_features))
        self.lora_B = nn.Parameter(torch.zeros(out_features, rank))
        self.scaling = alpha / rank  # python float fixed at construction; survives delattr (lora.py:24)
        self.lora_dropout = nn.Dropout(p=dropout) if dropout > 0.0 else (lambda x: x)
        nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B)

    def forward(self, x):
        result = F.linear(x, self.weight, self.bias)
        result = result + (self.lora_dropout(x) @ self.lora_A.T @ self.lora_B.T) * self.scaling
        return result


class RefLoRAConv2d(nn.Module):
    """loralib-0.1 ``Conv2d`` as adapted by ``modules/lora.py:15-16`` (k = kernel_size[0])."""

    def __init__(self, in_channels, out_channels, kernel_size, rank=4, alpha=1, dropout=0.0):
        super().__init__()
        k = kernel_size
        self.in_channels, self.out_channels, self.kernel_size, self.r = in_channels, out_channels, k, rank
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, k, k), requires_grad=False)
        self.bias = None
        self.lora_A = nn.Parameter(torch.zeros(rank * k, in_channels * k))
        self.lora_B = nn.Parameter(torch.zeros(out_channels * k, rank * k))
        self.scaling = alpha / rank
        nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B)

    def forward(self, x):
        w = self.weight + (self.lora_B @ self.lora_A).view(self.weight.shape) * self.scaling
        return F.conv2d(x, w, self.bias)  # loralib ctor saw no stride/padding -> defaults


def ref_get_lora(module: nn.Module, rank=4, alpha=1, dropout=0.0) -> nn.Module:
    """``modules/lora.py:12-27`` line for line, on the restated loralib classes."""
    if isinstance(module, nn.Linear):
        lora = RefLoRALinear(module.in_features, module.out_features, rank, alpha, dropout)
    elif isinstance(module, nn.Conv2d):
        lora = RefLoRAConv2d(module.in_channels, module.out_channels, module.kernel_size[0], rank, alpha, dropout)
    else:
        raise Exception("Unexpected module type")
    lora.weight = module.weight          # same Parameter object (lora.py:20)
    lora.bias = module.bias              # (lora.py:21)
    lora.lora_A.requires_grad = True
    lora.lora_B.requires_grad = True
    lora.register_buffer("lora_alpha", torch.tensor(alpha, dtype=torch.int32))
    return lora.to(module.weight.device)


# ----------------------------------------------------------------------------------------------
# Functional forms used by the parity tests (any dtype; fp64 gives the arbiter)
# ----------------------------------------------------------------------------------------------

def ref_lora_linear_fwd(x, w, bias, A, B, scaling):
    """``y = x W^T + b + s (x A^T) B^T`` with the reference's op order."""
    y = F.linear(x, w, bias)
    y = y + (x @ A.T @ B.T) * scaling
    return y


def ref_lora_linear_grads(x, w, bias, A, B, scaling, dy):
    """(y, dx, dA, dB) by autograd on the reference expression (W, b frozen: ``model.py:137``)."""
    x = x.detach().clone().requires_grad_(True)
    A = A.detach().clone().requires_grad_(True)
    B = B.detach().clone().requires_grad_(True)
    y = ref_lora_linear_fwd(x, w.detach(), None if bias is None else bias.detach(), A, B, scaling)
    y.backward(dy)
    return y.detach(), x.grad, A.grad, B.grad


def ref_lora_linear_grads_closed_form(x, w, A, B, scaling, dy):
    """SURVEY §8 a-1 closed forms, used to cross-check autograd:
    ``dX = dY W + s (dY B) A``; ``dA = s (dY B)^T X``; ``dB = s dY^T (X A^T)``."""
    x2, dy2 = x.reshape(-1, x.shape[-1]), dy.reshape(-1, dy.shape[-1])
    g = dy2 @ B
    dx = dy2 @ w + scaling * (g @ A)
    dA = scaling * (g.T @ x2)
    dB = scaling * (dy2.T @ (x2 @ A.T))
    return dx.reshape(x.shape), dA, dB


def ref_lora_weight_grads_chunked(x, A, B, scaling, dy, chunk=16384):
    """dA, dB of the closed forms above, accumulated over row chunks in fp64 -- the full reductions over M without ever
    holding ``[M, *]`` in double precision (the benchmark-size parity tests: M up to 98,304 tokens).  ``x`` / ``dy`` may live
    on any device and in any dtype; every chunk is brought to the CPU and widened before it is used."""
    A64, B64 = A.detach().double().cpu(), B.detach().double().cpu()
    dA = torch.zeros_like(A64)
    dB = torch.zeros_like(B64)
    M = x.shape[0]
    for i in range(0, M, chunk):
        xc = x[i:i + chunk].detach().double().cpu()
        dyc = dy[i:i + chunk].detach().double().cpu()
        g = dyc @ B64
        dA += scaling * (g.T @ xc)
        dB += scaling * (dyc.T @ (xc @ A64.T))
    return dA, dB


def ref_rows_matvec(t, vec, chunk=16384):
    """``t.double() @ vec`` on the CPU, row chunk by row chunk (``t`` any device / dtype)."""
    return torch.cat([t[i:i + chunk].detach().double().cpu() @ vec for i in range(0, t.shape[0], chunk)])


def ref_cols_vecmat(u, t, chunk=16384):
    """``u @ t.double()`` on the CPU, accumulated over row chunks."""
    out = torch.zeros(t.shape[1], dtype=torch.float64)
    for i in range(0, t.shape[0], chunk):
        out += u[i:i + chunk] @ t[i:i + chunk].detach().double().cpu()
    return out
