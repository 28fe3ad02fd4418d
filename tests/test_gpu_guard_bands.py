"""Out-of-bounds WRITE checks without a sanitizer: every output of a kernel call lies between two guard bands of a sentinel
pattern inside one larger allocation; the bands must be untouched afterwards.  Ragged sizes exercise the clipping paths
(row tails of the GEMM epilogue's transposed stores, partial slabs of the norm kernels)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GUARD = 4096          # bytes on each side
SENTINEL = 0x5A


class Guarded:
    """A tensor view with sentinel bytes before and after it."""

    def __init__(self, shape, dtype):
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        pad = (-nbytes) % 256
        self.raw = torch.full((GUARD + nbytes + pad + GUARD,), SENTINEL, dtype=torch.uint8, device=DEV)
        self.view = self.raw[GUARD:GUARD + nbytes].view(dtype).view(*shape)
        self.nbytes = nbytes

    def intact(self) -> bool:
        lo = self.raw[:GUARD]
        hi = self.raw[GUARD + self.nbytes:]
        return bool((lo == SENTINEL).all() and (hi == SENTINEL).all())


def st():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("M,K,N,R,G", [(600, 640, 640, 16, 1), (154, 768, 320, 16, 4), (1000, 320, 2568, 16, 1), (300, 320, 328, 32, 3),
                                       (257, 1280, 10240, 16, 1), (130, 64, 200, 64, 2)])
def test_gemm_outputs_stay_inside_their_buffers(sdt_lib, M, K, N, R, G):
    from scal_sdt_b200 import _lib
    lib = sdt_lib
    g = torch.Generator(device=DEV).manual_seed(M + N)
    x = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    ws = [(torch.randn(N, K, device=DEV, generator=g) / K ** 0.5).bfloat16() for _ in range(G)]
    bs = [torch.randn(N, device=DEV, generator=g) for _ in range(G)]
    As = [(torch.randn(R, K, device=DEV, generator=g) / K ** 0.5).bfloat16() for _ in range(G)]
    Bs = [(torch.randn(N, R, device=DEV, generator=g) * 0.1).bfloat16() for _ in range(G)]
    ys = [Guarded((M, N), torch.bfloat16) for _ in range(G)]
    ts = [Guarded((M, R), torch.bfloat16) for _ in range(G)]
    if G == 1:
        _lib.check(lib.sdt_lora_linear_fwd(x.data_ptr(), ws[0].data_ptr(), bs[0].data_ptr(), As[0].data_ptr(), Bs[0].data_ptr(), 0.5,
                                           ys[0].view.data_ptr(), ts[0].view.data_ptr(), M, K, N, R, 1, st()))
    else:
        probs = (_lib.LoraProblem * G)(*[_lib.LoraProblem(x.data_ptr(), ws[q].data_ptr(), bs[q].data_ptr(), As[q].data_ptr(),
                                                          Bs[q].data_ptr(), ys[q].view.data_ptr(), ts[q].view.data_ptr()) for q in range(G)])
        _lib.check(lib.sdt_lora_linear_fwd_group(ctypes.addressof(probs), G, 0.5, M, K, N, R, 1, st()))
    torch.cuda.synchronize()
    for q in range(G):
        assert ys[q].intact() and ts[q].intact(), f"guard band of problem {q} was written"
        ref = x.float() @ ws[q].float().t() + bs[q] + (0.5 * (x.float() @ As[q].float().t())).bfloat16().float() @ Bs[q].float().t()
        err = (ys[q].view.float() - ref).norm() / ref.norm()
        assert err < 2e-2, err
        assert torch.isfinite(ys[q].view.float()).all()          # every element was written (sentinel bf16 0x5A5A is finite, so compare)
        assert (ys[q].view.view(torch.int16) != 0x5A5A).float().mean() > 0.99


@pytest.mark.parametrize("M,K,N,R,G", [(600, 640, 640, 16, 3), (300, 320, 320, 32, 2)])
def test_summed_backward_outputs_stay_inside_their_buffers(sdt_lib, M, K, N, R, G):
    from scal_sdt_b200 import _lib
    lib = sdt_lib
    g = torch.Generator(device=DEV).manual_seed(M)
    x = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    dys = [torch.randn(M, N, device=DEV, generator=g).bfloat16() for _ in range(G)]
    wts = [(torch.randn(K, N, device=DEV, generator=g) / N ** 0.5).bfloat16() for _ in range(G)]
    Ats = [(torch.randn(K, R, device=DEV, generator=g) * 0.1).bfloat16() for _ in range(G)]
    Bts = [(torch.randn(R, N, device=DEV, generator=g) * 0.1).bfloat16() for _ in range(G)]
    tss = [torch.randn(M, R, device=DEV, generator=g).bfloat16() for _ in range(G)]
    dx = Guarded((M, K), torch.bfloat16)
    gws = [Guarded((M, R), torch.bfloat16) for _ in range(G)]
    dAs = [Guarded((R, K), torch.float32) for _ in range(G)]
    dBs = [Guarded((N, R), torch.float32) for _ in range(G)]
    for t in dAs + dBs:
        t.view.zero_()
    assert lib.sdt_lora_linear_bwd_group_supported(G, 1, M, K, N, R)
    probs = (_lib.LoraBwdProblem * G)(*[_lib.LoraBwdProblem(dys[q].data_ptr(), x.data_ptr(), wts[q].data_ptr(), Ats[q].data_ptr(),
                                                            Bts[q].data_ptr(), tss[q].data_ptr(), gws[q].view.data_ptr(),
                                                            dAs[q].view.data_ptr(), dBs[q].view.data_ptr()) for q in range(G)])
    _lib.check(lib.sdt_lora_linear_bwd_group(ctypes.addressof(probs), G, 0.5, dx.view.data_ptr(), M, K, N, R, R, 1, _lib.wgrad_workspace(), st()))
    torch.cuda.synchronize()
    assert dx.intact() and all(t.intact() for t in gws + dAs + dBs)
    ref = sum(dys[q].float() @ wts[q].float().t() + (0.5 * (dys[q].float() @ Bts[q].float().t())).bfloat16().float() @ Ats[q].float().t()
              for q in range(G))
    assert (dx.view.float() - ref).norm() / ref.norm() < 2e-2


@pytest.mark.parametrize("M,C", [(1001, 320), (77, 1280), (5, 2048), (33, 64)])
def test_layer_norm_outputs_stay_inside_their_buffers(sdt_lib, M, C):
    from scal_sdt_b200 import _lib
    lib = sdt_lib
    x = torch.randn(M, C, device=DEV).bfloat16()
    r = torch.randn(M, C, device=DEV).bfloat16()
    gam, bet = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    xs, y, dx = Guarded((M, C), torch.bfloat16), Guarded((M, C), torch.bfloat16), Guarded((M, C), torch.bfloat16)
    stats = Guarded((M, 2), torch.float32)
    _lib.check(lib.sdt_layer_norm_fwd(x.data_ptr(), r.data_ptr(), gam.data_ptr(), bet.data_ptr(), xs.view.data_ptr(), y.view.data_ptr(),
                                      stats.view.data_ptr(), M, C, 1e-5, st()))
    _lib.check(lib.sdt_layer_norm_bwd(xs.view.data_ptr(), y.view.data_ptr(), r.data_ptr(), gam.data_ptr(), stats.view.data_ptr(),
                                      dx.view.data_ptr(), M, C, st()))
    torch.cuda.synchronize()
    assert xs.intact() and y.intact() and dx.intact() and stats.intact()
    assert torch.equal(xs.view, x + r)


@pytest.mark.parametrize("B,HW,C", [(3, 35, 640), (2, 4096, 320), (5, 9, 2560), (1, 1, 64)])
def test_group_norm_and_residual_outputs_stay_inside_their_buffers(sdt_lib, B, HW, C):
    from scal_sdt_b200 import _lib
    lib = sdt_lib
    G = 32 if C % 32 == 0 and C // 32 >= 8 else 8
    x = torch.randn(B, HW, C, device=DEV).bfloat16()
    d = torch.randn(B, HW, C, device=DEV).bfloat16()
    cb = torch.randn(B, C, device=DEV).bfloat16()
    gam, bet = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    y, dx, out = (Guarded((B, HW, C), torch.bfloat16) for _ in range(3))
    n_ws = int(lib.sdt_group_norm_workspace_floats(B, G))
    stats, bst = Guarded((B, G, 2), torch.float32), Guarded((n_ws,), torch.float32)
    _lib.check(lib.sdt_group_norm_nhwc(x.data_ptr(), cb.data_ptr(), gam.data_ptr(), bet.data_ptr(), stats.view.data_ptr(),
                                       y.view.data_ptr(), B, HW, C, G, 1e-5, 1, bst.view.data_ptr(), n_ws, st()))
    _lib.check(lib.sdt_group_norm_nhwc_bwd(x.data_ptr(), cb.data_ptr(), d.data_ptr(), gam.data_ptr(), bet.data_ptr(),
                                           stats.view.data_ptr(), bst.view.data_ptr(), n_ws, dx.view.data_ptr(), B, HW, C, G, 1e-5, 1, st()))
    _lib.check(lib.sdt_residual_bias_add(x.data_ptr(), d.data_ptr(), gam.data_ptr(), out.view.data_ptr(), B * HW, C, st()))
    torch.cuda.synchronize()
    assert y.intact() and dx.intact() and out.intact() and stats.intact() and bst.intact()
    assert torch.isfinite(y.view.float()).all() and torch.isfinite(dx.view.float()).all()
    assert torch.equal(out.view, (x.float() + d.float() + 1.0).bfloat16())
