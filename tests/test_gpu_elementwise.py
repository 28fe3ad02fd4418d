"""GPU parity: noising / target / loss / EMA / AdamW kernels through the C ABI vs the CPU oracle.
fp32 results are bit-exact (the kernels reproduce torch's op-by-op rounding); bf16 likewise."""
import pytest
import torch

from oracle import diffusion_ref, ema_ref

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _inputs(B, C, H, W, dtype, seed=114514):
    g = torch.Generator().manual_seed(seed)
    x0 = torch.randn(B, C, H, W, generator=g).to(dtype)
    eps = torch.randn(B, C, H, W, generator=g).to(dtype)
    t = torch.randint(0, 1000, (B,), generator=g, dtype=torch.int64)
    return x0, eps, t


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(8, 4, 64, 64), (2, 4, 96, 128), (3, 4, 5, 7), (1, 4, 64, 64)])
@pytest.mark.parametrize("ptype", ["epsilon", "sample", "v"])
def test_noise_target_bit_exact(sdt_lib, dtype, shape, ptype):
    from scal_sdt_b200 import NoiseScheduler
    x0, eps, t = _inputs(*shape, dtype)
    sched = NoiseScheduler(prediction_type=ptype)
    noisy, target = sched.noise_and_target(x0.to(DEV), eps.to(DEV), t.to(DEV), check_range=True)
    ac = diffusion_ref.ref_alphas_cumprod()
    assert torch.equal(ac, sched.alphas_cumprod)
    ref_noisy = diffusion_ref.ref_add_noise(ac, x0, eps, t)
    ref_target = diffusion_ref.ref_target(ptype, ac, x0, eps, t)
    assert torch.equal(noisy.cpu(), ref_noisy)
    assert torch.equal(target.cpu(), ref_target)


def test_noise_target_extreme_timesteps_and_oob(sdt_lib):
    from scal_sdt_b200 import NoiseScheduler
    x0, eps, _ = _inputs(4, 4, 8, 8, torch.float32)
    t = torch.tensor([0, 999, 500, 1], dtype=torch.int64)
    sched = NoiseScheduler(prediction_type="v")
    noisy, v = sched.noise_and_target(x0.to(DEV), eps.to(DEV), t.to(DEV), check_range=True)
    ac = diffusion_ref.ref_alphas_cumprod()
    assert torch.equal(noisy.cpu(), diffusion_ref.ref_add_noise(ac, x0, eps, t))
    assert torch.equal(v.cpu(), diffusion_ref.ref_get_velocity(ac, x0, eps, t))
    with pytest.raises(IndexError):
        sched.noise_and_target(x0.to(DEV), eps.to(DEV), torch.tensor([0, 1000, 1, 2]).to(DEV), check_range=True)
    with pytest.raises(Exception, match="Unknown prediction type"):
        sched.noise_and_target(x0.to(DEV), eps.to(DEV), t.to(DEV), prediction_type="v_prediction")


@pytest.mark.parametrize("pdtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(8, 4, 64, 64), (4, 4, 96, 128), (2, 3, 5, 7)])
@pytest.mark.parametrize("prior", [False, True])
def test_mse_loss_and_grad(sdt_lib, pdtype, shape, prior):
    from scal_sdt_b200 import DenoiseLoss
    g = torch.Generator().manual_seed(1)
    pred = torch.randn(*shape, generator=g).to(pdtype)
    target = torch.randn(*shape, generator=g)
    crit = DenoiseLoss(DEV, prior_preservation=prior, prior_loss_weight=0.7)
    p = pred.to(DEV).requires_grad_(True)
    loss, elem = crit(p, target.to(DEV), want_elementwise=True)
    loss.backward()
    # oracle (fp64 arbiter for the scalar; fp32 elementwise bit-exact)
    pr = pred.clone().float().requires_grad_(True)
    ref_elem = diffusion_ref.ref_elementwise_loss(pr, target)
    ref_loss = diffusion_ref.ref_reduce_loss(ref_elem, prior, 0.7)
    ref_loss.backward()
    assert torch.equal(elem.cpu(), ref_elem.detach())
    assert abs(loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
    tol = 1e-5 if pdtype == torch.float32 else 2e-2
    gr, rr = p.grad.float().cpu(), pr.grad
    assert (gr - rr).norm() <= tol * rr.norm()
    crit.raise_if_nan()


def test_mse_nan_guard(sdt_lib):
    from scal_sdt_b200 import DenoiseLoss
    crit = DenoiseLoss(DEV)
    p = torch.randn(2, 4, 8, 8, device=DEV)
    p[1, 2, 3, 4] = float("nan")
    crit(p, torch.zeros_like(p))
    with pytest.raises(Exception, match="NaN element discovered in loss"):
        crit.raise_if_nan()


def _mlp(seed=0):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(33, 65), torch.nn.GELU(), torch.nn.Linear(65, 17), torch.nn.LayerNorm(17))


def test_ema_multi_tensor_bit_exact(sdt_lib):
    from scal_sdt_b200 import ExponentialMovingAverage
    m_ref, m_gpu = _mlp(), _mlp().to(DEV)
    ref = ema_ref.RefEMA(m_ref, 0.995)
    ema = ExponentialMovingAverage(m_gpu, 0.995)
    g = torch.Generator().manual_seed(3)
    for step in range(12):
        with torch.no_grad():
            for pr, pg in zip(m_ref.parameters(), m_gpu.parameters()):
                d = torch.randn(pr.shape, generator=g) * 0.1
                pr.add_(d)
                pg.add_(d.to(DEV))
        ref.update()
        ema.update()
        assert ema.num_updates == ref.num_updates
    for name, s in ref.shadow_params.items():
        assert torch.equal(ema.shadow_params[name].cpu(), s), name
    sd = ema.state_dict()
    assert set(sd) == {"decay", "num_updates", "shadow_params"} and sd["num_updates"] == 12


def test_ema_partially_frozen_and_flat_arena(sdt_lib):
    """LoRA-like: only some parameters train (the reference's own class raises KeyError here)."""
    from scal_sdt_b200 import ExponentialMovingAverage, ParamArena
    m_ref, m_gpu = _mlp(1), _mlp(1).to(DEV)
    for m in (m_ref, m_gpu):
        m[0].weight.requires_grad_(False)
        m[3].bias.requires_grad_(False)
    arena = ParamArena([{"params": [p for p in m_gpu.parameters() if p.requires_grad]}])
    ref = ema_ref.RefEMA(m_ref, 0.9)
    ema = ExponentialMovingAverage(m_gpu, 0.9)
    assert ema._flat is not None
    g = torch.Generator().manual_seed(4)
    for _ in range(5):
        with torch.no_grad():
            for pr, pg in zip(m_ref.parameters(), m_gpu.parameters()):
                if pr.requires_grad:
                    d = torch.randn(pr.shape, generator=g)
                    pr.add_(d)
                    pg.add_(d.to(DEV))
        ref.update()
        ema.update()
    assert set(ema.shadow_params) == set(ref.shadow_params)
    for name, s in ref.shadow_params.items():
        assert torch.equal(ema.shadow_params[name].cpu(), s), name
    with ema.average_parameters():
        for (n, p) in m_gpu.named_parameters():
            if n in ema.shadow_params:
                assert torch.equal(p, ema.shadow_params[n])
    assert arena.numel >= sum(p.numel() for p in m_gpu.parameters() if p.requires_grad)


def test_ema_large_flat_bf16_and_f32(sdt_lib):
    from scal_sdt_b200 import _lib
    lib = _lib.load()
    for dtype, code in ((torch.float32, 0), (torch.bfloat16, 1)):
        n = 3 * 1024 * 1024 + 5
        g = torch.Generator().manual_seed(7)
        s = torch.randn(n, generator=g).to(dtype)
        p = torch.randn(n, generator=g).to(dtype)
        sd, pd = s.to(DEV), p.to(DEV)
        omd = 1.0 - 0.995
        _lib.check(lib.sdt_ema_update_flat(sd.data_ptr(), pd.data_ptr(), n, omd, None, code, 0))
        torch.cuda.synchronize()
        tmp = s - p
        tmp.mul_(omd)
        s.sub_(tmp)
        assert torch.equal(sd.cpu(), s)


def test_flat_adamw_matches_torch(sdt_lib):
    from scal_sdt_b200 import FlatAdamW, ParamArena
    m_ref, m_gpu = _mlp(2), _mlp(2).to(DEV)
    groups_ref = [{"params": list(m_ref[0].parameters()), "lr": 5e-4, "weight_decay": 2e-2},
                  {"params": list(m_ref[2].parameters()) + list(m_ref[3].parameters()), "lr": 5e-3, "weight_decay": 2e-3}]
    groups_gpu = [{"params": list(m_gpu[0].parameters()), "lr": 5e-4, "weight_decay": 2e-2},
                  {"params": list(m_gpu[2].parameters()) + list(m_gpu[3].parameters()), "lr": 5e-3, "weight_decay": 2e-3}]
    opt_ref = torch.optim.AdamW(groups_ref, lr=1e-3, betas=(0.9, 0.999), eps=1e-7, weight_decay=1e-2)
    arena = ParamArena(groups_gpu)
    opt = FlatAdamW(arena, lr=1e-3, betas=(0.9, 0.999), eps=1e-7, weight_decay=1e-2)
    g = torch.Generator().manual_seed(5)
    for _ in range(10):
        for pr, pg in zip(m_ref.parameters(), m_gpu.parameters()):
            gr = torch.randn(pr.shape, generator=g)
            pr.grad = gr.clone()
            pg.grad.copy_(gr.to(DEV))
        opt_ref.step()
        opt.step()
    for (n, pr), pg in zip(m_ref.named_parameters(), m_gpu.parameters()):
        err = (pg.cpu() - pr).abs().max().item()
        assert err <= 2e-6 * max(1.0, pr.abs().max().item()), (n, err)


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 1e-2), (torch.float32, 1e-5)])
def test_geglu_forward_backward(sdt_lib, dtype, tol):
    from scal_sdt_b200.fused import geglu
    g = torch.Generator().manual_seed(9)
    proj = torch.randn(3, 37, 2 * 1280, generator=g).to(dtype)
    dout = torch.randn(3, 37, 1280, generator=g).to(dtype)
    pr = proj.double().requires_grad_(True)
    h, gate = pr.chunk(2, dim=-1)
    ref = h * torch.nn.functional.gelu(gate)
    ref.backward(dout.double())
    po = proj.to(DEV).requires_grad_(True)
    out = geglu(po)
    out.backward(dout.to(DEV))
    assert out.shape == ref.shape and out.dtype == dtype
    assert (out.double().cpu() - ref).norm() <= tol * ref.norm()
    assert (po.grad.double().cpu() - pr.grad).norm() <= tol * pr.grad.norm()


@pytest.mark.parametrize("C,G,H,W,silu", [(320, 32, 16, 16, True), (960, 32, 8, 12, True), (640, 32, 7, 5, False),
                                          (2560, 32, 4, 4, True), (1280, 32, 16, 16, False), (320, 32, 64, 64, True),
                                          (1920, 32, 32, 32, True), (64, 8, 3, 3, False)])
def test_group_norm_nhwc_forward_backward(sdt_lib, C, G, H, W, silu):
    from scal_sdt_b200.fused import group_norm_act, group_norm_nhwc_supported
    torch.manual_seed(C + H)
    norm = torch.nn.GroupNorm(G, C, eps=1e-5)
    with torch.no_grad():
        norm.weight.normal_(1.0, 0.3)
        norm.bias.normal_(0.0, 0.3)
    x = (torch.randn(3, C, H, W) * 1.7 + 0.4).bfloat16()
    dout = torch.randn(3, C, H, W).bfloat16()
    # reference: fp64 on the same bf16-rounded inputs and parameters
    ref_norm = torch.nn.GroupNorm(G, C, eps=1e-5).double()
    with torch.no_grad():
        ref_norm.weight.copy_(norm.weight.bfloat16().double())
        ref_norm.bias.copy_(norm.bias.bfloat16().double())
    xr = x.double().requires_grad_(True)
    yr = ref_norm(xr)
    if silu:
        yr = torch.nn.functional.silu(yr)
    yr.backward(dout.double())
    gn = norm.to(DEV).to(torch.bfloat16).requires_grad_(False)
    xo = x.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    assert group_norm_nhwc_supported(gn, xo)
    yo = group_norm_act(gn, xo, silu)
    assert yo.is_contiguous(memory_format=torch.channels_last) and yo.dtype == torch.bfloat16
    yo.backward(dout.to(DEV))
    assert (yo.double().cpu() - yr).norm() <= 1e-2 * yr.norm()
    assert (xo.grad.double().cpu() - xr.grad).norm() <= 1e-2 * xr.grad.norm()
    # the statistics are summed in a fixed order (per-thread values in thread order, per-CTA partials in CTA order): run to run
    # the kernels give the same bits, like torch's native GroupNorm (round 1 used float atomics)
    for _ in range(3):
        x2 = x.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        y2 = group_norm_act(gn, x2, silu)
        y2.backward(dout.to(DEV))
        assert torch.equal(y2, yo) and torch.equal(x2.grad, xo.grad)


@pytest.mark.parametrize("C,G,H,W,silu", [(320, 32, 16, 16, True), (1280, 32, 8, 8, True), (640, 32, 7, 5, False)])
def test_group_norm_with_folded_time_embedding_bias(sdt_lib, C, G, H, W, silu):
    """``silu(norm(h + temb[:, :, None, None]))`` with the broadcast add folded into the norm kernels: output, dX and the
    bias gradient against fp64 autograd on the same bf16 operands."""
    from scal_sdt_b200.fused import group_norm_act
    torch.manual_seed(C + W)
    norm = torch.nn.GroupNorm(G, C, eps=1e-5)
    with torch.no_grad():
        norm.weight.normal_(1.0, 0.3)
        norm.bias.normal_(0.0, 0.3)
    x = (torch.randn(3, C, H, W) * 1.3 - 0.2).bfloat16()
    tb = (torch.randn(3, C) * 0.8).bfloat16()
    dout = torch.randn(3, C, H, W).bfloat16()
    ref_norm = torch.nn.GroupNorm(G, C, eps=1e-5).double()
    with torch.no_grad():
        ref_norm.weight.copy_(norm.weight.bfloat16().double())
        ref_norm.bias.copy_(norm.bias.bfloat16().double())
    xr, tr = x.double().requires_grad_(True), tb.double().requires_grad_(True)
    yr = ref_norm(xr + tr[:, :, None, None])
    if silu:
        yr = torch.nn.functional.silu(yr)
    yr.backward(dout.double())
    gn = norm.to(DEV).to(torch.bfloat16).requires_grad_(False)
    xo = x.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    to = tb.to(DEV).requires_grad_(True)
    yo = group_norm_act(gn, xo, silu, chan_bias=to)
    yo.backward(dout.to(DEV))
    assert (yo.double().cpu() - yr).norm() <= 1e-2 * yr.norm()
    assert (xo.grad.double().cpu() - xr.grad).norm() <= 1e-2 * xr.grad.norm()
    assert (to.grad.double().cpu() - tr.grad).norm() <= 2e-2 * tr.grad.norm()


def test_residual_bias_add_kernel(sdt_lib):
    from scal_sdt_b200.fused import residual_bias_add
    g = torch.Generator().manual_seed(4)
    a = torch.randn(3, 320, 9, 7, generator=g).bfloat16().to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    b = torch.randn(3, 320, 9, 7, generator=g).bfloat16().to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    bias = torch.randn(320, generator=g).to(DEV)
    out = residual_bias_add(a, b, bias)
    ref = a.float() + b.float() + bias[None, :, None, None]
    assert out.dtype == torch.bfloat16 and out.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(out, ref.bfloat16())                         # one rounding of the exact f32 sum
    dout = torch.randn_like(out)
    out.backward(dout)
    assert torch.equal(a.grad, dout) and torch.equal(b.grad, dout)


@pytest.mark.parametrize("cin,cout", [(320, 320), (320, 640)])
def test_resnet_block_with_folded_biases_matches_plain_evaluation(sdt_lib, cin, cout):
    """ResnetBlock2D with frozen parameters (conv biases + time embedding folded into norm2 / the residual add) against the
    plain module-by-module evaluation in fp64 on the same bf16-rounded parameters and inputs."""
    import copy

    import torch.nn.functional as F
    from scal_sdt_b200.unet import ResnetBlock2D
    torch.manual_seed(cin + cout)
    blk = ResnetBlock2D(cin, cout, 1280, 32)
    with torch.no_grad():
        for p in blk.parameters():
            p.copy_(p.bfloat16().float())
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.1).copy_(p.bfloat16().float())
    ref = copy.deepcopy(blk).double()
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, cin, 16, 16, generator=g).bfloat16()
    temb = torch.randn(2, 1280, generator=g).bfloat16()
    dout = torch.randn(2, cout, 16, 16, generator=g).bfloat16()
    xr = x.double().requires_grad_(True)
    h = ref.conv1(F.silu(ref.norm1(xr)))
    h = h + ref.time_emb_proj(F.silu(temb.double()))[:, :, None, None]
    h = ref.conv2(F.silu(ref.norm2(h)))
    yr = (ref.conv_shortcut(xr) if ref.conv_shortcut is not None else xr) + h
    yr.backward(dout.double())
    gpu = blk.to(DEV).to(torch.bfloat16).to(memory_format=torch.channels_last).requires_grad_(False)
    assert gpu._frozen_biases() is not None
    xo = x.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    yo = gpu(xo, temb.to(DEV))
    yo.backward(dout.to(DEV))
    assert (yo.double().cpu() - yr).norm() <= 2e-2 * yr.norm(), (yo.double().cpu() - yr).norm() / yr.norm()
    assert (xo.grad.double().cpu() - xr.grad).norm() <= 3e-2 * xr.grad.norm(), (xo.grad.double().cpu() - xr.grad).norm() / xr.grad.norm()
