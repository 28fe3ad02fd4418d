"""GPU parity of the LoRA projection (K1 forward, K2 backward) through the C ABI and through ``get_lora``.

Tolerances are the ones north_star states: 1e-5 relative for fp32, 2e-2 for bf16, on outputs and LoRA gradients,
measured as ||ours - ref||_F / ||ref||_F against the oracle (``oracle/lora_ref.py``) evaluated in fp64/fp32 on the
same (bf16-rounded, for the bf16 case) inputs.  lora_B is non-zero, otherwise dA and the LoRA part of dX vanish."""
import pytest
import torch
from torch import nn

from oracle import lora_ref

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-300)).item()


def make_pair(kind, cin, cout, rank, alpha, bias, seed, dtype):
    """(reference module on CPU, ours on GPU) with identical frozen weights and non-zero LoRA factors."""
    torch.manual_seed(seed)
    base = nn.Linear(cin, cout, bias=bias) if kind == "linear" else nn.Conv2d(cin, cout, 1, bias=bias)
    ref = lora_ref.ref_get_lora(base, rank, alpha)
    with torch.no_grad():
        ref.lora_B.normal_(0, 0.2)
    import copy
    from scal_sdt_b200 import get_lora
    base_gpu = copy.deepcopy(base).to(DEV)
    base_gpu.requires_grad_(False)
    ours = get_lora(base_gpu, rank, alpha)
    with torch.no_grad():
        ours.lora_A.copy_(ref.lora_A)
        ours.lora_B.copy_(ref.lora_B)
    if dtype == torch.bfloat16:      # compare on identical (bf16-representable) operands
        with torch.no_grad():
            for m in (ref, ours):
                m.weight.copy_(m.weight.bfloat16().float())
                m.lora_A.copy_(m.lora_A.bfloat16().float())
                m.lora_B.copy_(m.lora_B.bfloat16().float())
    return ref.double(), ours


CASES = [
    # kind, shape of x, cin, cout, rank, alpha, bias
    ("linear", (2, 77, 768), 768, 320, 4, 1, False),       # cross-attention to_k, SD1.5, rank 4 (cfg1)
    ("linear", (2, 256, 320), 320, 320, 16, 1, True),      # to_out.0 rank 16 (cfg2)
    ("linear", (1, 1000, 640), 640, 5120, 16, 8, True),    # ff.net.0.proj, ragged M
    ("linear", (2, 64, 1280), 1280, 1280, 64, 64, False),  # rank 64 (cfg3)
    ("linear", (3, 50, 1024), 1024, 640, 32, 16, False),   # SD2.x-shaped context, BN=128 path
    ("conv", (2, 320, 16, 16), 320, 320, 16, 1, True),     # proj_in 1x1 conv
]


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 2e-2), (torch.float32, 1e-5)])
@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}-{c[2]}x{c[3]}-r{c[4]}" for c in CASES])
def test_get_lora_forward_backward(sdt_lib, case, dtype, tol):
    kind, xshape, cin, cout, rank, alpha, bias = case
    ref, ours = make_pair(kind, cin, cout, rank, alpha, bias, 7, dtype)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(*xshape, generator=g).to(dtype).float()
    xr = x.double().requires_grad_(True)
    yr = ref(xr)
    dy = torch.randn(yr.shape, generator=g).to(dtype).float()
    yr.backward(dy.double())
    xo = x.to(DEV).to(dtype).requires_grad_(True)
    yo = ours(xo)
    assert yo.shape == yr.shape and yo.dtype == dtype
    yo.backward(dy.to(DEV).to(dtype))
    assert rel(yo, yr) <= tol, ("y", rel(yo, yr))
    assert rel(xo.grad, xr.grad) <= tol, ("dx", rel(xo.grad, xr.grad))
    assert rel(ours.lora_A.grad, ref.lora_A.grad) <= tol, ("dA", rel(ours.lora_A.grad, ref.lora_A.grad))
    assert rel(ours.lora_B.grad, ref.lora_B.grad) <= tol, ("dB", rel(ours.lora_B.grad, ref.lora_B.grad))
    assert ours.weight.grad is None and (ours.bias is None or ours.bias.grad is None)


def test_get_lora_contract(sdt_lib):
    """Attribute / state-dict contract of modules/lora.py:12-27."""
    from scal_sdt_b200 import get_lora
    base = nn.Linear(64, 32).to(DEV)
    m = get_lora(base, rank=4, alpha=2)
    assert m.weight is base.weight and m.bias is base.bias
    assert m.lora_A.shape == (4, 64) and m.lora_B.shape == (32, 4)
    assert m.lora_A.requires_grad and m.lora_B.requires_grad and m.lora_A.dtype == torch.float32
    assert m.lora_alpha.dtype == torch.int32 and int(m.lora_alpha) == 2 and m.scaling == 0.5
    assert torch.count_nonzero(m.lora_B) == 0 and m.lora_A.abs().max() <= 1 / 8 + 1e-6
    assert set(m.state_dict()) == {"weight", "bias", "lora_A", "lora_B", "lora_alpha"}
    assert m.lora_A.device == base.weight.device
    with pytest.raises(Exception, match="Unexpected module type"):
        get_lora(nn.LayerNorm(8))


def test_zero_init_lora_equals_base(sdt_lib):
    """With lora_B = 0 (the reference's init) the injected module reproduces the frozen projection."""
    from scal_sdt_b200 import get_lora
    torch.manual_seed(0)
    base = nn.Linear(320, 640).to(DEV).to(torch.bfloat16)
    m = get_lora(base, rank=16, alpha=1)
    x = torch.randn(4, 100, 320, device=DEV, dtype=torch.bfloat16)
    y = m(x)
    ref = torch.nn.functional.linear(x.float(), base.weight.float(), base.bias.float())
    assert rel(y, ref) <= 4e-3      # one bf16 rounding of the output


def test_arena_direct_gradients_and_pack(sdt_lib):
    """LoraArena: backward accumulates straight into the flat gradient arena; pack() refreshes bf16 operands."""
    from scal_sdt_b200 import FlatAdamW, LoraArena, config_module
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(64, 128), nn.GELU(), nn.Linear(128, 64)).to(DEV)
    groups = config_module(net, [{"index": ["0", "2"], "lora": {"rank": 8, "alpha": 4}, "optimizer": {"lr": 1e-2}}])
    arena = LoraArena(net, groups)
    with torch.no_grad():
        for _, m in arena.sites:
            m.lora_B.normal_(0, 0.1)
    arena.pack()
    opt = FlatAdamW(arena, lr=1e-3)
    x = torch.randn(32, 64, device=DEV, dtype=torch.bfloat16)
    losses = []
    for _ in range(3):
        opt.zero_grad()
        y = net(x)
        loss = y.float().pow(2).mean()
        loss.backward()
        assert net[0].lora_A.grad.data_ptr() == arena.grad_view(net[0].lora_A).data_ptr()
        assert arena.grads.abs().sum() > 0
        opt.step()
        arena.pack()
        losses.append(loss.item())
    assert losses[2] < losses[0]


RAGGED = [(1, 64, 64), (7, 320, 320), (129, 72, 40), (255, 320, 640), (257, 640, 320), (616, 768, 1280), (1000, 8, 8),
          (300, 1280, 24)]


@pytest.mark.parametrize("M,K,N", RAGGED, ids=[f"{m}x{k}x{n}" for m, k, n in RAGGED])
@pytest.mark.parametrize("rank", [4, 16])
def test_ragged_shapes_bf16(sdt_lib, M, K, N, rank):
    """Token counts / widths that do not fill the 128 x 160 x 64 tiles (TMA zero-fill on loads, clipping on stores)."""
    ref, ours = make_pair("linear", K, N, rank, 2 * rank, True, 3, torch.bfloat16)
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g).bfloat16().float()
    dy = torch.randn(M, N, generator=g).bfloat16().float()
    xr = x.double().requires_grad_(True)
    yr = ref(xr)
    yr.backward(dy.double())
    xo = x.to(DEV).bfloat16().requires_grad_(True)
    yo = ours(xo)
    yo.backward(dy.to(DEV).bfloat16())
    for name, a, b in (("y", yo, yr), ("dx", xo.grad, xr.grad), ("dA", ours.lora_A.grad, ref.lora_A.grad),
                       ("dB", ours.lora_B.grad, ref.lora_B.grad)):
        assert rel(a, b) <= 2e-2, (name, rel(a, b))


def test_empty_batch(sdt_lib):
    from scal_sdt_b200 import get_lora
    m = get_lora(nn.Linear(64, 32).to(DEV), rank=4)
    y = m(torch.zeros(0, 64, device=DEV, dtype=torch.bfloat16))
    assert y.shape == (0, 32)
    y.sum().backward()
    assert m.lora_A.grad is not None and float(m.lora_A.grad.abs().sum()) == 0.0


def test_largest_bucket_token_count_linearity(sdt_lib):
    """cfg4's largest bucket at batch 8: M = 8 * 128 * 96 = 98,304 tokens.  Size-independent checks: the projection is
    linear in x (f(x1 + x2) - f(0) = (f(x1) - f(0)) + (f(x2) - f(0))) and a random row sample matches the oracle."""
    torch.manual_seed(0)
    M, K, N, r = 98304, 320, 320, 16
    ref, ours = make_pair("linear", K, N, r, 16, True, 9, torch.bfloat16)
    g = torch.Generator(device=DEV).manual_seed(1)
    x1 = (torch.randn(M, K, device=DEV, generator=g) * 0.5).bfloat16()
    x2 = (torch.randn(M, K, device=DEV, generator=g) * 0.5).bfloat16()
    with torch.no_grad():
        f0 = ours(torch.zeros(1, K, device=DEV, dtype=torch.bfloat16)).float()
        y1, y2, y12 = ours(x1).float(), ours(x2).float(), ours((x1.float() + x2.float()).bfloat16()).float()
    lin = (y12 - f0) - ((y1 - f0) + (y2 - f0))
    assert lin.norm() <= 2e-2 * y12.norm()
    rows = torch.randint(0, M, (257,), generator=torch.Generator().manual_seed(2))
    yr = ref(x1[rows.to(DEV)].cpu().double())
    assert rel(y1[rows.to(DEV)], yr) <= 2e-2


GROUP_CASES = [
    # x shape, cin, cout, rank, alpha, bias, number of projections
    ((2, 256, 320), 320, 320, 16, 1, False, 3),     # attn1 q/k/v, level 0 width -> single-CTA kernel (K < 512)
    ((2, 300, 640), 640, 640, 16, 8, False, 3),     # q/k/v, ragged M -> CTA-pair kernel (K >= 512)
    ((2, 77, 768), 768, 320, 4, 1, False, 4),       # k/v of two cross-attentions on one text context (M = 154 < 256)
    ((8, 77, 768), 768, 1280, 64, 64, True, 2),     # k/v, rank 64, with bias, M = 616 -> pair kernel
    ((1, 130, 1024), 1024, 640, 32, 16, False, 3),  # SD2.x context width, BN = 128 tiles
    ((4, 128, 640), 640, 640, 32, 16, False, 3),    # rank 32: three 32-column rank accumulators fill TMEM exactly (summed dX)
    ((2, 256, 320), 320, 640, 16, 1, False, 2),     # two sources, N != K
    ((8, 4096, 320), 320, 320, 16, 1, False, 3),    # cfg2 level-0 q/k/v at full size (M = 32768)
]


@pytest.mark.parametrize("case", GROUP_CASES, ids=[f"{c[1]}x{c[2]}-r{c[3]}-g{c[6]}" for c in GROUP_CASES])
def test_grouped_projection_matches_oracle_and_single_launches(sdt_lib, case):
    """``project_group`` (one launch for G same-shape sites on a shared input) against the oracle, and bit-for-bit against
    the per-site launches it replaces; gradients flow to the shared input and to every site's lora_A / lora_B."""
    from scal_sdt_b200.lora import groupable, project_group
    xshape, cin, cout, rank, alpha, bias, G = case
    pairs = [make_pair("linear", cin, cout, rank, alpha, bias, 100 + g, torch.bfloat16) for g in range(G)]
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(*xshape, generator=gen).bfloat16().float()
    dys = [torch.randn(*xshape[:-1], cout, generator=gen).bfloat16().float() for _ in range(G)]
    # oracle
    xr = x.double().requires_grad_(True)
    yrs = [ref(xr) for ref, _ in pairs]
    torch.autograd.backward(yrs, [d.double() for d in dys])
    # grouped launch
    mods = [ours for _, ours in pairs]
    xo = x.to(DEV).bfloat16().requires_grad_(True)
    assert groupable(mods, xo.reshape(-1, cin))
    yos = project_group(mods, xo)
    torch.autograd.backward(yos, [d.to(DEV).bfloat16() for d in dys])
    for g in range(G):
        assert rel(yos[g], yrs[g]) <= 2e-2, ("y", g, rel(yos[g], yrs[g]))
        assert rel(mods[g].lora_A.grad, pairs[g][0].lora_A.grad) <= 2e-2, ("dA", g)
        assert rel(mods[g].lora_B.grad, pairs[g][0].lora_B.grad) <= 2e-2, ("dB", g)
    assert rel(xo.grad, xr.grad) <= 2e-2, ("dx", rel(xo.grad, xr.grad))
    # the same sites one launch each: identical tiles, identical arithmetic order
    xs = x.to(DEV).bfloat16()
    for g in range(G):
        assert torch.equal(mods[g](xs), yos[g]), f"grouped launch differs from the single launch (site {g})"


def test_group_falls_back_to_per_site_launches_on_mixed_shapes(sdt_lib):
    from scal_sdt_b200 import get_lora
    from scal_sdt_b200.lora import groupable, project_group
    a = get_lora(nn.Linear(320, 320, bias=False).to(DEV).bfloat16(), 16, 1)
    b = get_lora(nn.Linear(320, 640, bias=False).to(DEV).bfloat16(), 16, 1)
    x = torch.randn(4, 64, 320, device=DEV, dtype=torch.bfloat16)
    assert not groupable([a, b], x.reshape(-1, 320))
    ya, yb = project_group([a, b], x)
    assert torch.equal(ya, a(x)) and torch.equal(yb, b(x))


@pytest.mark.parametrize("G,cout,rank", [(4, 320, 4), (2, 1280, 16), (3, 640, 64)])
def test_grouped_projection_backward_without_input_gradient(sdt_lib, G, cout, rank):
    """to_k / to_v of cross-attentions read the text context, which needs no gradient: the group's rank projections
    G = s dY B run as one launch; dA / dB must match the oracle."""
    from scal_sdt_b200.lora import project_group
    pairs = [make_pair("linear", 768, cout, rank, 2, False, 300 + g, torch.bfloat16) for g in range(G)]
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(8, 77, 768, generator=gen).bfloat16().float()
    dys = [torch.randn(8, 77, cout, generator=gen).bfloat16().float() for _ in range(G)]
    yrs = [ref(x.double()) for ref, _ in pairs]
    torch.autograd.backward(yrs, [d.double() for d in dys])
    mods = [ours for _, ours in pairs]
    yos = project_group(mods, x.to(DEV).bfloat16())            # no requires_grad on the input
    torch.autograd.backward(yos, [d.to(DEV).bfloat16() for d in dys])
    for g in range(G):
        assert rel(yos[g], yrs[g]) <= 2e-2
        assert rel(mods[g].lora_A.grad, pairs[g][0].lora_A.grad) <= 2e-2, ("dA", g, rel(mods[g].lora_A.grad, pairs[g][0].lora_A.grad))
        assert rel(mods[g].lora_B.grad, pairs[g][0].lora_B.grad) <= 2e-2, ("dB", g)


@pytest.mark.parametrize("M,K,N,rank", [(32768, 320, 320, 16), (2048, 1280, 5120, 64), (1000, 640, 640, 4), (616, 768, 320, 16)])
def test_lora_gradients_are_bit_reproducible(sdt_lib, M, K, N, rank):
    """dA / dB are reductions over tens of thousands of token rows split over ~148 CTAs.  The partial sums are combined in a
    fixed order (per-slice partials + last-arriver reduction inside the one launch), so repeated backward passes give
    bit-identical gradients -- as the reference's torch path does; and accumulation into an existing gradient still works."""
    ref, ours = make_pair("linear", K, N, rank, 2 * rank, True, 3, torch.bfloat16)
    g = torch.Generator(device=DEV).manual_seed(M + K + N)
    x = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    dy = torch.randn(M, N, device=DEV, generator=g).bfloat16()
    runs = []
    for _ in range(4):
        ours.lora_A.grad = ours.lora_B.grad = None
        xo = x.clone().requires_grad_(True)
        ours(xo).backward(dy)
        runs.append((ours.lora_A.grad.clone(), ours.lora_B.grad.clone(), xo.grad.clone()))
    for dA, dB, dx in runs[1:]:
        assert torch.equal(dA, runs[0][0]) and torch.equal(dB, runs[0][1]) and torch.equal(dx, runs[0][2])
    # second backward without clearing: the kernels accumulate (+=) into the existing gradient
    xo = x.clone().requires_grad_(True)
    ours(xo).backward(dy)
    assert rel(ours.lora_A.grad, 2 * runs[0][0]) <= 1e-6 and rel(ours.lora_B.grad, 2 * runs[0][1]) <= 1e-6
    # and the values are still the oracle's
    dA_ref, dB_ref = lora_ref.ref_lora_weight_grads_chunked(x, ref.lora_A, ref.lora_B, ref.scaling, dy)
    assert rel(runs[0][0], dA_ref) <= 2e-2 and rel(runs[0][1], dB_ref) <= 2e-2


F16_CASES = [
    ("linear", (2, 77, 768), 768, 320, 4, 1, False),        # single-CTA kernel, rank 4
    ("linear", (8, 1024, 640), 640, 640, 16, 16, True),     # CTA-pair kernel with bias (bias hi/lo pair in fp16)
    ("linear", (1, 1000, 640), 640, 5120, 16, 8, True),     # 224-wide tiles, ragged M
    ("linear", (2, 300, 1280), 1280, 1280, 64, 64, False),  # rank 64
    ("conv", (2, 320, 16, 16), 320, 320, 16, 1, True),      # proj_in 1x1 conv
]


@pytest.mark.parametrize("case", F16_CASES, ids=[f"{c[0]}-{c[2]}x{c[3]}-r{c[4]}" for c in F16_CASES])
def test_get_lora_forward_backward_fp16(sdt_lib, case):
    """The reference's stock configs train with ``trainer.precision: 16`` (configs/lora.yaml:50): IEEE half operands on the same
    tcgen05 kernels (instruction-descriptor format field + fp16 conversions).  Bound: 2e-2 as for bf16 -- and, half having three
    more mantissa bits than bf16, the measured error must also be well below the bf16 run's."""
    kind, xshape, cin, cout, rank, alpha, bias = case
    errs = {}
    for dtype in (torch.float16, torch.bfloat16):
        ref, ours = make_pair(kind, cin, cout, rank, alpha, bias, 7, torch.bfloat16)   # operands exact in both 16-bit formats
        with torch.no_grad():                          # ... if they also fit fp16's range / precision: round through both
            for m in (ref, ours):
                for p in (m.weight, m.lora_A, m.lora_B):
                    p.copy_(p.float().half().bfloat16().half().float().to(p.dtype))
        g = torch.Generator().manual_seed(11)
        x = torch.randn(*xshape, generator=g).bfloat16().half().float()
        xr = x.double().requires_grad_(True)
        yr = ref(xr)
        dy = torch.randn(yr.shape, generator=g).bfloat16().half().float()
        yr.backward(dy.double())
        xo = x.to(DEV).to(dtype).requires_grad_(True)
        yo = ours(xo)
        assert yo.dtype == dtype
        yo.backward(dy.to(DEV).to(dtype))
        errs[dtype] = (rel(yo, yr), rel(xo.grad, xr.grad), rel(ours.lora_A.grad, ref.lora_A.grad), rel(ours.lora_B.grad, ref.lora_B.grad))
    for e in errs[torch.float16]:
        assert e <= 2e-2, errs
    assert errs[torch.float16][0] < 0.5 * errs[torch.bfloat16][0], errs        # y: one rounding of the output dominates
    assert errs[torch.float16][1] < 0.5 * errs[torch.bfloat16][1], errs        # dx likewise


def test_fp16_grouped_qkv_and_autocast(sdt_lib):
    """q / k / v as one fp16 launch (+ summed-source dX), and fp32 masters driven through torch.autocast(float16)."""
    from scal_sdt_b200.lora import groupable, project_group
    pairs = [make_pair("linear", 640, 640, 16, 16, False, 100 + g, torch.bfloat16) for g in range(3)]
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(2, 300, 640, generator=gen).bfloat16().float()
    dys = [torch.randn(2, 300, 640, generator=gen).bfloat16().float() for _ in range(3)]
    xr = x.double().requires_grad_(True)
    yrs = [ref(xr) for ref, _ in pairs]
    torch.autograd.backward(yrs, [d.double() for d in dys])
    mods = [ours for _, ours in pairs]
    xo = x.to(DEV).half().requires_grad_(True)
    assert groupable(mods, xo.reshape(-1, 640))
    yos = project_group(mods, xo)
    torch.autograd.backward(yos, [d.to(DEV).half() for d in dys])
    for g in range(3):
        assert yos[g].dtype == torch.float16 and rel(yos[g], yrs[g]) <= 2e-2
        assert rel(mods[g].lora_A.grad, pairs[g][0].lora_A.grad) <= 2e-2 and rel(mods[g].lora_B.grad, pairs[g][0].lora_B.grad) <= 2e-2
    assert rel(xo.grad, xr.grad) <= 2e-2
    # autocast: fp32 input and fp32 frozen weight, fp16 compute
    ref, ours = pairs[0]
    x32 = x.to(DEV)
    with torch.autocast("cuda", dtype=torch.float16):
        y = ours(x32)
    assert y.dtype == torch.float16 and rel(y, yrs[0]) <= 2e-2


def test_batched_weight_gradients_of_mixed_shapes_match_per_site_launches(sdt_lib):
    """``sdt_lora_wgrad_batch``: the dA / dB reductions of a transformer block's worth of sites -- different token counts
    (8192 image tokens, 616 text tokens), widths and bias-ness, one padded rank -- in ONE launch: equal to the per-site
    launches up to f32 summation order (the token range is cut into different slices), bit-reproducible itself, and equal to
    the oracle."""
    from scal_sdt_b200 import LoraArena, config_module
    from scal_sdt_b200.lora import deferred_wgrad

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.q, self.out = nn.Linear(320, 320, bias=False), nn.Linear(320, 320)
            self.k = nn.Linear(768, 320, bias=False)
            self.up, self.down = nn.Linear(320, 2560), nn.Linear(1280, 320)
    torch.manual_seed(0)
    net = Block().to(DEV).bfloat16()
    groups = config_module(net, [{"index": ["q", "out", "k", "up", "down"], "lora": {"rank": 16, "alpha": 16}}])
    arena = LoraArena(net, groups)
    with torch.no_grad():
        for _, m in arena.sites:
            m.lora_B.normal_(0, 0.1)
    arena.pack()
    g = torch.Generator(device=DEV).manual_seed(1)
    xs = {"q": torch.randn(8192, 320, device=DEV, generator=g).bfloat16().requires_grad_(True),
          "out": torch.randn(8192, 320, device=DEV, generator=g).bfloat16().requires_grad_(True),
          "k": torch.randn(616, 768, device=DEV, generator=g).bfloat16(),
          "up": torch.randn(8192, 320, device=DEV, generator=g).bfloat16().requires_grad_(True),
          "down": torch.randn(8192, 1280, device=DEV, generator=g).bfloat16().requires_grad_(True)}
    dys = {n: torch.randn(xs[n].shape[0], getattr(net, n).out_features, device=DEV, generator=g).bfloat16() for n in xs}

    def backward_all():
        arena.zero_grad()
        for x in xs.values():
            x.grad = None
        ys = [getattr(net, n)(xs[n]) for n in xs]
        torch.autograd.backward(ys, [dys[n] for n in xs])
        return arena.grads.clone(), {n: (x.grad.clone() if x.grad is not None else None) for n, x in xs.items()}
    per_site, dx_a = backward_all()
    with deferred_wgrad() as q:
        arena.zero_grad()
        ys = [getattr(net, n)(xs[n]) for n in xs]
        torch.autograd.backward(ys, [dys[n] for n in xs])
        assert len(q.items) == 5 and float(arena.grads.abs().sum()) == 0.0       # nothing reduced yet
    batched = arena.grads.clone()
    assert (batched - per_site).norm() <= 1e-5 * per_site.norm()
    with deferred_wgrad():
        arena.zero_grad()
        ys = [getattr(net, n)(xs[n]) for n in xs]
        torch.autograd.backward(ys, [dys[n] for n in xs])
    assert torch.equal(arena.grads, batched)                                      # run to run: bit-identical
    for n in xs:
        m = getattr(net, n)
        dA_ref, dB_ref = lora_ref.ref_lora_weight_grads_chunked(xs[n].detach(), m.lora_A, m.lora_B, m.scaling, dys[n])
        assert rel(m.lora_A.grad, dA_ref) <= 2e-2 and rel(m.lora_B.grad, dB_ref) <= 2e-2, n


GEGLU_CASES = [(32768, 320, 1280, 16), (8192, 640, 2560, 16), (2048, 1280, 5120, 16), (1000, 640, 2560, 64), (300, 320, 1280, 4)]


@pytest.mark.parametrize("M,K,I,rank", GEGLU_CASES, ids=[f"{m}x{k}x2*{i}-r{r}" for m, k, i, r in GEGLU_CASES])
def test_geglu_epilogue_equals_projection_then_geglu(sdt_lib, M, K, I, rank, monkeypatch):
    """ff.net.0.proj with the GEGLU activation formed in the GEMM's epilogue: bit-identical to the projection
    launch followed by the GEGLU kernel (same rounded proj values go into the same arithmetic), equal to the fp64 oracle
    ``h * gelu(gate)`` of the reference expression within the bf16 bound, same gradients."""
    from scal_sdt_b200.fused import geglu
    from scal_sdt_b200.lora import geglu_projection, geglu_projection_supported
    ref, ours = make_pair("linear", K, 2 * I, rank, rank, True, 5, torch.bfloat16)
    g = torch.Generator(device=DEV).manual_seed(M + I)
    x = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    dact = torch.randn(M, I, device=DEV, generator=g).bfloat16()
    assert not geglu_projection_supported(ours, x)          # opt-in: measured slower than the two launches (profiles/)
    monkeypatch.setenv("SDT_FUSED_GEGLU", "1")
    assert geglu_projection_supported(ours, x)
    xa = x.clone().requires_grad_(True)
    act = geglu_projection(ours, xa)
    act.backward(dact)
    gA, gB, gx = ours.lora_A.grad.clone(), ours.lora_B.grad.clone(), xa.grad.clone()
    ours.lora_A.grad = ours.lora_B.grad = None
    xb = x.clone().requires_grad_(True)
    act2 = geglu(ours(xb))
    act2.backward(dact)
    assert torch.equal(act, act2)                                            # fused epilogue == two launches, to the bit
    assert torch.equal(gx, xb.grad) and torch.equal(gA, ours.lora_A.grad) and torch.equal(gB, ours.lora_B.grad)
    # oracle on a row sample (all columns): the reference expression in fp64
    rows = torch.randperm(M, generator=torch.Generator().manual_seed(1))[:384]
    xr = x[rows.to(DEV)].double().cpu()
    proj = ref(xr)
    h, gate = proj.chunk(2, dim=-1)
    act_ref = h * torch.nn.functional.gelu(gate)
    assert rel(act[rows.to(DEV)], act_ref) <= 2e-2, rel(act[rows.to(DEV)], act_ref)


@pytest.mark.parametrize("M,K,N,rank,dtype", [(32768, 1280, 320, 16, torch.bfloat16), (8192, 640, 640, 16, torch.bfloat16),
                                              (1000, 2560, 640, 64, torch.bfloat16), (100, 320, 320, 4, torch.bfloat16),
                                              (2048, 5120, 1280, 16, torch.float16)])
def test_residual_added_in_the_epilogue(sdt_lib, M, K, N, rank, dtype):
    """``proj(x) + residual`` (ff.net.2 / proj_out and the block's residual stream) with the add in the projection's epilogue:
    the bits of the projection launch followed by torch's add, gradient of the residual = gradient of the output, and the oracle
    expression within the bf16 / fp16 bound."""
    ref, ours = make_pair("linear", K, N, rank, rank, True, 4, torch.bfloat16)
    g = torch.Generator(device=DEV).manual_seed(M + N)
    x = torch.randn(M, K, device=DEV, generator=g).to(dtype)
    res = torch.randn(M, N, device=DEV, generator=g).to(dtype)
    dy = torch.randn(M, N, device=DEV, generator=g).to(dtype)
    xa, ra = x.clone().requires_grad_(True), res.clone().requires_grad_(True)
    ya = ours(xa, residual=ra)
    ya.backward(dy)
    gA, gB = ours.lora_A.grad.clone(), ours.lora_B.grad.clone()
    ours.lora_A.grad = ours.lora_B.grad = None
    xb, rb = x.clone().requires_grad_(True), res.clone().requires_grad_(True)
    yb = ours(xb) + rb
    yb.backward(dy)
    assert torch.equal(ya, yb)
    assert torch.equal(xa.grad, xb.grad) and torch.equal(ra.grad, rb.grad) and torch.equal(ra.grad, dy)
    assert torch.equal(gA, ours.lora_A.grad) and torch.equal(gB, ours.lora_B.grad)
    rows = torch.randperm(M, generator=torch.Generator().manual_seed(2))[:256]
    yr = ref(x[rows.to(DEV)].double().cpu()) + res[rows.to(DEV)].double().cpu()
    assert rel(ya[rows.to(DEV)], yr) <= 2e-2


@pytest.mark.parametrize("rank,bias,dtype", [(16, False, torch.bfloat16), (64, False, torch.bfloat16), (4, True, torch.float16)])
def test_multi_width_projection_of_the_text_context(sdt_lib, rank, bias, dtype):
    """to_k / to_v of every cross-attention on ONE text context as the work items of one launch (sdt_lora_linear_fwd_multi):
    bit-identical to the per-site launches, gradients equal to the oracle's, the context itself gets no gradient."""
    from scal_sdt_b200.lora import multi_projectable, project_multi
    widths = [320, 320, 640, 640, 1280, 1280, 1280, 320, 640, 1280, 1280]
    pairs = [make_pair("linear", 768, n, rank, rank, bias, 500 + i, torch.bfloat16) for i, n in enumerate(widths)]
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(8, 77, 768, generator=gen).bfloat16().float()
    dys = [torch.randn(8, 77, n, generator=gen).bfloat16().float() for n in widths]
    yrs = [ref(x.double()) for ref, _ in pairs]
    torch.autograd.backward(yrs, [d.double() for d in dys])
    mods = [ours for _, ours in pairs]
    xo = x.to(DEV).to(dtype)
    assert multi_projectable(mods, xo.reshape(-1, 768))
    yos = project_multi(mods, xo)
    torch.autograd.backward(yos, [d.to(DEV).to(dtype) for d in dys])
    for g, n in enumerate(widths):
        assert yos[g].shape == (8, 77, n) and yos[g].dtype == dtype
        assert rel(yos[g], yrs[g]) <= 2e-2, ("y", g, rel(yos[g], yrs[g]))
        assert torch.equal(mods[g](xo), yos[g]), f"multi-width launch differs from the single launch (site {g})"
        assert rel(mods[g].lora_A.grad, pairs[g][0].lora_A.grad) <= 2e-2 and rel(mods[g].lora_B.grad, pairs[g][0].lora_B.grad) <= 2e-2, g
    assert not multi_projectable(mods, xo.reshape(-1, 768).clone().requires_grad_(True))     # inputs that need dX go per site


@pytest.mark.parametrize("M,K,N,rank,bias", [(8192, 1280, 640, 16, True), (15300, 2560, 320, 16, False), (8192 + 77, 1344, 640, 64, True),
                                             (16384, 1280, 320, 32, True)])
def test_double_tiles_equal_single_tiles(sdt_lib, M, K, N, rank, bias):
    """K >= 1280 and N % 320 == 0 with enough row tiles selects the double-tile kernel (the two column tiles of a work item in one
    joint K loop, X landed once).  Same UMMAs in the same order per tile: the output and the saved rank intermediate must be
    bit-identical to the single-tile kernel (``sdt_debug_set(31, 1)`` switches double tiles off), forward and backward, for ragged
    M, a K that is not a multiple of the 64-deep k-block, and every padded rank."""
    from scal_sdt_b200 import _lib, get_lora
    lib = _lib.load()
    torch.manual_seed(M + K)
    base = nn.Linear(K, N, bias=bias).to(DEV).requires_grad_(False)
    mod = get_lora(base, rank=rank, alpha=rank)
    with torch.no_grad():
        mod.lora_B.normal_(0.0, 0.1)
    x = torch.randn(M, K, device=DEV).bfloat16()
    dy = torch.randn(M, N, device=DEV).bfloat16()

    def run():
        xi = x.clone().requires_grad_(True)
        mod.lora_A.grad = mod.lora_B.grad = None
        y = mod(xi)
        y.backward(dy)
        return y.detach().clone(), xi.grad.clone(), mod.lora_A.grad.clone(), mod.lora_B.grad.clone()
    try:
        lib.sdt_debug_set(31, 1)
        single = run()
        lib.sdt_debug_set(31, 0)
        double = run()
    finally:
        lib.sdt_debug_set(31, 0)
    for name, a, b in zip(("y", "dx", "dA", "dB"), single, double):
        assert torch.equal(a, b), f"{name}: double tiles differ from single tiles"
    # and the pair is right: a stratified sample of rows against the fp64 oracle
    rows = torch.arange(0, M, 97, device=DEV)
    w, b, A, B = (t.detach().double() for t in (mod.weight, mod.bias if bias else torch.zeros(N, device=DEV), mod.lora_A, mod.lora_B))
    xr = x[rows].double()
    ref = xr @ w.T + b + mod.scaling * (xr @ A.T) @ B.T
    assert rel(double[0][rows], ref) <= 2e-2
