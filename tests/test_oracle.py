"""CPU tier: the oracle restatements against the golden vectors produced by the reference's own source
(``oracle/make_golden.py``) and against closed-form identities where the arithmetic lives in absent third-party
packages (loralib 0.1, diffusers DDIMScheduler -- parity unpinned by the reference, see oracle/__init__.py)."""
import json
import math
from pathlib import Path

import pytest
import torch
from torch import nn

from oracle import diffusion_ref, ema_ref, lora_ref

GOLDEN = Path(__file__).parent / "golden"


def _mlp():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(33, 65), nn.GELU(), nn.Linear(65, 17), nn.LayerNorm(17))


def test_ema_restatement_matches_reference_golden():
    gold = torch.load(GOLDEN / "ema_reference.pt")
    m = _mlp()
    for k, v in m.state_dict().items():
        assert torch.equal(v, gold["init"][k])
    ema = ema_ref.RefEMA(m, gold["decay"])
    for step, deltas in enumerate(gold["deltas"]):
        with torch.no_grad():
            for p, d in zip(m.parameters(), deltas):
                p.add_(d)
        ema.update()
        assert ema_ref.RefEMA.decay_at(gold["decay"], ema.num_updates) == gold["decays"][step]
    assert ema.num_updates == gold["num_updates"]
    for k, v in gold["shadow"].items():
        assert torch.equal(ema.shadow_params[k], v), k          # bit-exact


def test_ema_warmup_schedule():
    assert [ema_ref.RefEMA.decay_at(0.995, n) for n in (1, 2, 3)] == [2 / 11, 3 / 12, 4 / 13]
    assert ema_ref.RefEMA.decay_at(0.995, 10 ** 6) == 0.995


def test_reference_ema_crashes_when_partially_frozen():
    """SURVEY fact 6, recorded from the reference's own class; the restatement walks shadow keys instead."""
    rec = json.loads((GOLDEN / "ema_reference_partial_freeze.json").read_text())
    assert rec["reference_raises"] is not None and "KeyError" in rec["reference_raises"]
    m = nn.Sequential(nn.Linear(4, 4), nn.Linear(4, 4))
    m[0].weight.requires_grad_(False)
    e = ema_ref.RefEMA(m, 0.9)
    e.update()
    assert "0.weight" not in e.shadow_params and "0.bias" in e.shadow_params


def test_ema_state_dict_roundtrip_and_average_parameters():
    m = _mlp()
    e = ema_ref.RefEMA(m, 0.9)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(1.0)
    e.update()
    sd = e.state_dict()
    e2 = ema_ref.RefEMA(_mlp(), 0.5)
    e2.load_state_dict(sd)
    assert e2.decay == 0.9 and e2.num_updates == 1
    before = [p.clone() for p in m.parameters()]
    with e.average_parameters():
        for (n, p) in m.named_parameters():
            assert torch.equal(p, e.shadow_params[n])
    for p, b in zip(m.parameters(), before):
        assert torch.equal(p, b)
    with pytest.raises(ValueError):
        ema_ref.RefEMA(m, 1.5)


# ---- LoRA (loralib 0.1 semantics) ----------------------------------------------------------------------
def test_lora_linear_merged_weight_equivalence_fp64():
    torch.manual_seed(1)
    base = nn.Linear(48, 24).double()
    m = lora_ref.ref_get_lora(base, rank=4, alpha=2).double()
    with torch.no_grad():
        m.lora_B.normal_()
    x = torch.randn(5, 7, 48, dtype=torch.float64)
    merged = base.weight + (m.lora_B @ m.lora_A) * (2 / 4)
    assert torch.allclose(m(x), torch.nn.functional.linear(x, merged, base.bias), rtol=1e-12, atol=1e-12)


def test_lora_closed_form_gradients_match_autograd():
    torch.manual_seed(2)
    M, K, N, r, s = 40, 32, 24, 4, 0.25
    x, w = torch.randn(M, K, dtype=torch.float64), torch.randn(N, K, dtype=torch.float64)
    A, B = torch.randn(r, K, dtype=torch.float64), torch.randn(N, r, dtype=torch.float64)
    dy = torch.randn(M, N, dtype=torch.float64)
    _, dx, dA, dB = lora_ref.ref_lora_linear_grads(x, w, None, A, B, s, dy)
    dx2, dA2, dB2 = lora_ref.ref_lora_linear_grads_closed_form(x, w, A, B, s, dy)
    for a, b in ((dx, dx2), (dA, dA2), (dB, dB2)):
        assert torch.allclose(a, b, rtol=1e-11, atol=1e-11)


def test_lora_conv1x1_equals_per_pixel_linear():
    torch.manual_seed(3)
    base = nn.Conv2d(16, 8, 1).double()
    m = lora_ref.ref_get_lora(base, rank=4, alpha=4).double()
    with torch.no_grad():
        m.lora_B.normal_()
    x = torch.randn(2, 16, 5, 6, dtype=torch.float64)
    y = m(x)
    tokens = x.permute(0, 2, 3, 1).reshape(-1, 16)
    ref = lora_ref.ref_lora_linear_fwd(tokens, base.weight.view(8, 16), base.bias, m.lora_A, m.lora_B, m.scaling)
    assert torch.allclose(y.permute(0, 2, 3, 1).reshape(-1, 8), ref, rtol=1e-12, atol=1e-12)


def test_get_lora_contract_on_reference_restatement():
    base = nn.Linear(64, 32)
    m = lora_ref.ref_get_lora(base, rank=4, alpha=1)
    assert m.weight is base.weight and m.bias is base.bias
    assert m.lora_A.shape == (4, 64) and m.lora_B.shape == (32, 4) and torch.count_nonzero(m.lora_B) == 0
    assert m.lora_alpha.dtype == torch.int32 and m.scaling == 0.25
    assert m.lora_A.abs().max() <= 1 / math.sqrt(64) + 1e-7        # kaiming_uniform(a=sqrt 5): U(+-1/sqrt(in))
    with pytest.raises(Exception, match="Unexpected module type"):
        lora_ref.ref_get_lora(nn.LayerNorm(4))


# ---- DDPM schedule / noising / loss ----------------------------------------------------------------------
def test_alphas_cumprod_properties():
    ac = diffusion_ref.ref_alphas_cumprod()
    assert ac.shape == (1000,) and ac.dtype == torch.float32
    assert torch.all(ac[1:] < ac[:-1]) and 0 < ac[-1] < ac[0] < 1
    assert abs(ac[0].item() - (1 - 0.00085)) < 1e-7
    assert abs(ac[-1].item() - 0.0046601) < 2e-6          # SD1.x terminal alpha_bar


def test_noising_identities():
    ac = diffusion_ref.ref_alphas_cumprod().double()
    g = torch.Generator().manual_seed(0)
    x0 = torch.randn(4, 4, 8, 8, generator=g, dtype=torch.float64)
    eps = torch.randn(4, 4, 8, 8, generator=g, dtype=torch.float64)
    t = torch.tensor([0, 10, 500, 999])
    noisy = diffusion_ref.ref_add_noise(ac, x0, eps, t)
    v = diffusion_ref.ref_get_velocity(ac, x0, eps, t)
    a = ac[t].sqrt().view(-1, 1, 1, 1)
    b = (1 - ac[t]).sqrt().view(-1, 1, 1, 1)
    # (noisy, v) is a rotation of (x0, eps): inverting it recovers both
    assert torch.allclose(a * noisy - b * v, x0, atol=1e-12)
    assert torch.allclose(b * noisy + a * v, eps, atol=1e-12)
    assert diffusion_ref.ref_target("epsilon", ac, x0, eps, t) is eps
    assert diffusion_ref.ref_target("sample", ac, x0, eps, t) is x0
    with pytest.raises(Exception, match="Unknown prediction type"):
        diffusion_ref.ref_target("v_prediction", ac, x0, eps, t)


def test_prior_preservation_reduction():
    g = torch.Generator().manual_seed(0)
    loss = torch.rand(6, 4, 3, 3, generator=g)
    r = diffusion_ref.ref_reduce_loss(loss, True, 0.3)
    assert torch.allclose(r, loss[:3].mean() + 0.3 * loss[3:].mean())
    assert torch.allclose(diffusion_ref.ref_reduce_loss(loss), loss.mean())


def test_chunked_weight_grads_and_projection_identities():
    """The helpers the benchmark-size GPU tests lean on: chunked dA/dB == autograd, and the +-1 projection closed forms
    (Y v, u^T Y, dX v, u^T dX) hold for the oracle's own outputs to fp64 round-off."""
    from oracle import lora_ref
    g = torch.Generator().manual_seed(0)
    M, K, N, r, s = 300, 40, 56, 8, 0.75
    x, dy = torch.randn(M, K, generator=g).double(), torch.randn(M, N, generator=g).double()
    w, b = torch.randn(N, K, generator=g).double(), torch.randn(N, generator=g).double()
    A, B = torch.randn(r, K, generator=g).double(), torch.randn(N, r, generator=g).double()
    y, dx, dA, dB = lora_ref.ref_lora_linear_grads(x, w, b, A, B, s, dy)
    dA_c, dB_c = lora_ref.ref_lora_weight_grads_chunked(x.float(), A, B, s, dy.float(), chunk=64)
    assert torch.allclose(dA_c, dA, rtol=1e-6, atol=1e-6) and torch.allclose(dB_c, dB, rtol=1e-6, atol=1e-6)
    vN, vK, uM = torch.randn(N, generator=g).double(), torch.randn(K, generator=g).double(), torch.randn(M, generator=g).double()
    yv = lora_ref.ref_rows_matvec(x, w.T @ vN, chunk=64) + float(b @ vN) + s * lora_ref.ref_rows_matvec(x, A.T @ (B.T @ vN), chunk=64)
    assert torch.allclose(yv, y @ vN, rtol=1e-10, atol=1e-10)
    ux = lora_ref.ref_cols_vecmat(uM, x, chunk=64)
    assert torch.allclose(ux @ w.T + uM.sum() * b + s * ((ux @ A.T) @ B.T), uM @ y, rtol=1e-10, atol=1e-10)
    dxv = lora_ref.ref_rows_matvec(dy, w @ vK, chunk=64) + s * lora_ref.ref_rows_matvec(dy, B @ (A @ vK), chunk=64)
    assert torch.allclose(dxv, dx @ vK, rtol=1e-10, atol=1e-10)
    udy = lora_ref.ref_cols_vecmat(uM, dy, chunk=64)
    assert torch.allclose(udy @ w + s * ((udy @ B) @ A), uM @ dx, rtol=1e-10, atol=1e-10)
