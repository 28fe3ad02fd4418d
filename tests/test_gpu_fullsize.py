"""Benchmark-size GPU parity of the fused LoRA projection: every (M, K, N) the BASELINE configs launch.

cfg2's twelve shapes plus the text-context k/v shapes at rank 16, the same at rank 64 (cfg3) and the largest bucket of cfg4
(M = 8 * 128 * 96 = 98,304 tokens) with the level-0 widths.  These are the sizes the kernel dispatch keys on (224-wide tiles,
one work item per tile, K >= 2048 schedule), so small-shape parity says nothing about them.

Every launch goes through the C ABI (``get_lora`` -> ctypes -> ``sdt_lora_linear_fwd/bwd``).  The fp64 oracle
(``oracle/lora_ref.py``) is applied at full size where that is a reduction (dA, dB: all M rows), on a stratified row sample
for Y and dX (at least one row of EVERY 128-row tile, all columns), and -- so that no row and no column of the full
outputs goes unchecked -- through random +-1 projections of the whole tensors (Y v, u^T Y, dX v, u^T dX against their
closed forms, O(M (K + N)) on the CPU).

Bounds: Frobenius-relative 2e-2 (north_star, bf16) AND element-wise |ours - ref| <= 2e-2 * max(|ref|, rms of the row).
"""
import pytest
import torch
from torch import nn

from oracle import lora_ref

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 2e-2

CFG2 = [(32768, 320, 320), (32768, 320, 2560), (32768, 1280, 320), (8192, 640, 640), (8192, 640, 5120), (8192, 2560, 640),
        (2048, 1280, 1280), (2048, 1280, 10240), (2048, 5120, 1280), (512, 1280, 1280), (512, 1280, 10240), (512, 5120, 1280),
        (616, 768, 320), (616, 768, 640), (616, 768, 1280)]
CFG4_LARGEST_BUCKET = [(98304, 320, 320), (98304, 320, 2560), (98304, 1280, 320)]
CASES = [(m, k, n, 16) for m, k, n in CFG2] + [(m, k, n, 64) for m, k, n in CFG2] + [(m, k, n, 16) for m, k, n in CFG4_LARGEST_BUCKET]


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-300)).item()


def elementwise_excess(ours, ref):
    """max over elements of |ours - ref| / max(|ref|, rms(row of ref)); must stay <= TOL."""
    ours, ref = ours.detach().double().cpu(), ref.detach().double().cpu()
    scale = torch.maximum(ref.abs(), ref.pow(2).mean(dim=-1, keepdim=True).sqrt().expand_as(ref))
    return ((ours - ref).abs() / (scale + 1e-300)).max().item()


def stratified_rows(M, gen, budget=512):
    """>= 1 row of every 128-row tile (both CTAs of a 256-row pair tile), at most ~budget rows... but never fewer than one
    per tile."""
    tiles = (M + 127) // 128
    per = max(1, budget // tiles)
    rows = []
    for t in range(tiles):
        lo, hi = t * 128, min(M, t * 128 + 128)
        rows.append(lo + torch.randperm(hi - lo, generator=gen)[:per])
    rows = torch.cat(rows)
    return torch.unique(torch.cat([rows, torch.tensor([0, M - 1])]))


def build_site(K, N, r, bias, seed):
    """(frozen W [N,K], bias, A, B) bf16-representable fp32 on the CPU + our module on the GPU with the same values."""
    from scal_sdt_b200 import get_lora
    g = torch.Generator().manual_seed(seed)
    base = nn.Linear(K, N, bias=bias)
    with torch.no_grad():
        base.weight.copy_(((torch.rand(N, K, generator=g) * 2 - 1) / K ** 0.5).bfloat16().float())
        if bias:
            base.bias.copy_(torch.randn(N, generator=g) * 0.1)
    base = base.to(DEV).requires_grad_(False)
    ours = get_lora(base, rank=r, alpha=r)                       # scaling 1: the LoRA branch is as large as the base branch
    with torch.no_grad():
        ours.lora_A.copy_(ours.lora_A.bfloat16().float())
        ours.lora_B.copy_((torch.randn(N, r, generator=g) * 0.2).bfloat16().float().to(DEV))
    return ours


def oracle_operands(ours):
    w = ours.weight.detach().double().cpu()
    b = None if ours.bias is None else ours.bias.detach().double().cpu()
    return w, b, ours.lora_A.detach().double().cpu(), ours.lora_B.detach().double().cpu()


@pytest.mark.parametrize("M,K,N,r", CASES, ids=[f"{m}x{k}x{n}-r{r}" for m, k, n, r in CASES])
def test_site_at_benchmark_size(sdt_lib, M, K, N, r):
    bias = M != 616                                              # to_k / to_v on the text context have no bias
    ours = build_site(K, N, r, bias, seed=M + K + N + r)
    s = ours.scaling
    w, b, A, B = oracle_operands(ours)
    gd = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn(M, K, device=DEV, generator=gd).bfloat16().requires_grad_(True)
    dy = torch.randn(M, N, device=DEV, generator=gd).bfloat16()
    y = ours(x)
    y.backward(dy)
    dx = x.grad
    torch.cuda.synchronize()

    gen = torch.Generator().manual_seed(5)
    rows = stratified_rows(M, gen)
    xr, dyr = x.detach()[rows.to(DEV)].double().cpu(), dy[rows.to(DEV)].double().cpu()
    y_ref, dx_ref, _, _ = lora_ref.ref_lora_linear_grads(xr, w, b, A, B, s, dyr)
    for name, got, ref in (("y", y.detach()[rows.to(DEV)], y_ref), ("dx", dx[rows.to(DEV)], dx_ref)):
        assert rel(got, ref) <= TOL, (name, "sampled rows", rel(got, ref))
        assert elementwise_excess(got, ref) <= TOL, (name, "element-wise", elementwise_excess(got, ref))

    # LoRA gradients: full reductions over all M rows
    dA_ref, dB_ref = lora_ref.ref_lora_weight_grads_chunked(x.detach(), A, B, s, dy)
    for name, got, ref in (("dA", ours.lora_A.grad, dA_ref), ("dB", ours.lora_B.grad, dB_ref)):
        assert rel(got, ref) <= TOL, (name, rel(got, ref))
        assert elementwise_excess(got, ref) <= TOL, (name, "element-wise", elementwise_excess(got, ref))

    # every row and every column of the full Y and dX, through random +-1 projections
    vN = (torch.randint(0, 2, (N,), generator=gen) * 2 - 1).double()
    vK = (torch.randint(0, 2, (K,), generator=gen) * 2 - 1).double()
    uM = (torch.randint(0, 2, (M,), generator=gen) * 2 - 1).double()
    xd, dyd = x.detach(), dy
    # Y v = X (W^T v) + b.v + s (X A^T)(B^T v)              u^T Y = (u^T X) W^T + (sum u) b + s ((u^T X) A^T) B^T
    t_rows = lora_ref.ref_rows_matvec(xd, A.T @ (B.T @ vN))      # X (A^T (B^T v)) -- the rank path collapsed onto one vector
    yv_ref = lora_ref.ref_rows_matvec(xd, w.T @ vN) + (0.0 if b is None else float(b @ vN)) + s * t_rows
    ux = lora_ref.ref_cols_vecmat(uM, xd)
    uy_ref = ux @ w.T + (0.0 if b is None else uM.sum() * b) + s * ((ux @ A.T) @ B.T)
    # dX v = dY (W v) + s (dY B)(A v)                         u^T dX = (u^T dY) W + s ((u^T dY) B) A
    dxv_ref = lora_ref.ref_rows_matvec(dyd, w @ vK) + s * lora_ref.ref_rows_matvec(dyd, B @ (A @ vK))
    udy = lora_ref.ref_cols_vecmat(uM, dyd)
    udx_ref = udy @ w + s * ((udy @ B) @ A)
    y_row_norm = lora_ref.ref_rows_matvec(y.detach().float().pow(2), torch.ones(N, dtype=torch.float64)).sqrt()
    dx_row_norm = lora_ref.ref_rows_matvec(dx.float().pow(2), torch.ones(K, dtype=torch.float64)).sqrt()
    y_col_norm = lora_ref.ref_cols_vecmat(torch.ones(M, dtype=torch.float64), y.detach().float().pow(2)).sqrt()
    dx_col_norm = lora_ref.ref_cols_vecmat(torch.ones(M, dtype=torch.float64), dx.float().pow(2)).sqrt()
    checks = (("Y v (all rows)", lora_ref.ref_rows_matvec(y.detach(), vN), yv_ref, y_row_norm),
              ("u^T Y (all columns)", lora_ref.ref_cols_vecmat(uM, y.detach()), uy_ref, y_col_norm),
              ("dX v (all rows)", lora_ref.ref_rows_matvec(dx, vK), dxv_ref, dx_row_norm),
              ("u^T dX (all columns)", lora_ref.ref_cols_vecmat(uM, dx), udx_ref, dx_col_norm))
    for name, got, ref, norm in checks:
        # a projection sums the rounding noise of a whole row / column with random signs; a wrong tile shifts it by O(norm)
        worst = ((got - ref).abs() / (norm + 1e-300)).max().item()
        assert worst <= TOL, (name, worst)


GROUPS = [(32768, 320, 16), (8192, 640, 16), (2048, 1280, 16), (8192, 640, 64), (2048, 1280, 64), (512, 1280, 64)]


@pytest.mark.parametrize("M,C,r", GROUPS, ids=[f"qkv-{m}x{c}-r{r}" for m, c, r in GROUPS])
def test_grouped_qkv_at_benchmark_size(sdt_lib, M, C, r):
    """attn1 to_q / to_k / to_v as ONE forward launch and ONE summed-source dX launch (rank 64: per-site dX launches)."""
    from scal_sdt_b200.lora import groupable, project_group
    mods = [build_site(C, C, r, False, seed=10 * C + r + j) for j in range(3)]
    gd = torch.Generator(device=DEV).manual_seed(4)
    x = torch.randn(M, C, device=DEV, generator=gd).bfloat16().requires_grad_(True)
    dys = [torch.randn(M, C, device=DEV, generator=gd).bfloat16() for _ in range(3)]
    assert groupable(mods, x)
    ys = project_group(mods, x)
    torch.autograd.backward(ys, dys)
    torch.cuda.synchronize()
    gen = torch.Generator().manual_seed(6)
    rows = stratified_rows(M, gen)
    xr = x.detach()[rows.to(DEV)].double().cpu()
    dx_ref = torch.zeros_like(xr)
    for j, m in enumerate(mods):
        w, b, A, B = oracle_operands(m)
        y_ref, dxj, _, _ = lora_ref.ref_lora_linear_grads(xr, w, b, A, B, m.scaling, dys[j][rows.to(DEV)].double().cpu())
        dx_ref += dxj
        assert rel(ys[j].detach()[rows.to(DEV)], y_ref) <= TOL, ("y", j)
        assert elementwise_excess(ys[j].detach()[rows.to(DEV)], y_ref) <= TOL, ("y element-wise", j)
        dA_ref, dB_ref = lora_ref.ref_lora_weight_grads_chunked(x.detach(), A, B, m.scaling, dys[j])
        assert rel(m.lora_A.grad, dA_ref) <= TOL and elementwise_excess(m.lora_A.grad, dA_ref) <= TOL, ("dA", j)
        assert rel(m.lora_B.grad, dB_ref) <= TOL and elementwise_excess(m.lora_B.grad, dB_ref) <= TOL, ("dB", j)
    got = x.grad[rows.to(DEV)]
    assert rel(got, dx_ref) <= TOL, ("dx", rel(got, dx_ref))
    assert elementwise_excess(got, dx_ref) <= TOL, ("dx element-wise", elementwise_excess(got, dx_ref))
