"""LoRA module injection -- drop-in for the reference's ``modules/lora.py``.

``get_lora(module, rank, alpha, dropout)`` keeps the reference signature and the attribute contract
(``modules/lora.py:12-27``): the returned module aliases the source module's frozen ``weight`` / ``bias``
Parameter objects, owns trainable fp32 ``lora_A [r,in]`` (kaiming-uniform, a=sqrt 5) and ``lora_B [out,r]``
(zeros), carries ``lora_alpha`` as an int32 buffer and lives on ``module.weight.device``.  ``scaling`` is
``alpha / rank`` fixed at construction (loralib 0.1).

The arithmetic ``y = x W^T + b + (alpha/r) (x A^T) B^T`` and its backward (dX, dA, dB; W, b frozen) run in
``libsdt_b200.so``: one fused tcgen05/TMEM/TMA kernel per direction for bf16, FFMA kernels for fp32.  There is
no torch fallback.
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Optional

import torch
from torch import nn

from . import _lib
from ._lib import PackSite, SdtError

MAX_TC_RANK = 64

# bench.py sets this to a list to time every projection launch with CUDA events on the launching stream.  One record per
# lora_gemm* kernel launch: ("fwd", M, K, N, R, G, start_event, end_event) / ("bwd", M, K, N, R, G, need_dx, start, end),
# G = number of same-shape projections the launch computes (1 unless grouped)
PROFILE = None


def _ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def padded_rank(rank: int) -> int:
    """Rank as seen by the tensor-core kernels (UMMA K granule is 16; supported 16 / 32 / 64)."""
    for r in (16, 32, 64):
        if rank <= r:
            return r
    raise SdtError(f"LoRA rank {rank} > {MAX_TC_RANK} is not supported by the bf16 tensor-core path")


def _pack_sites(sites: list[PackSite], max_elems: int, device, dtype=torch.bfloat16) -> torch.Tensor:
    arr = (PackSite * len(sites))(*sites)
    host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
    dev = host.to(device, non_blocking=False)
    _lib.check(_lib.load().sdt_lora_pack(dev.data_ptr(), len(sites), max_elems, _lib.dtype_code(dtype), _lib.stream_ptr()),
               "sdt_lora_pack")
    # `dev` must outlive the launch: the caching allocator keeps stream order for us
    return dev


class PackedOperands:
    """16-bit (bf16 or fp16) tensor-core operand layouts of one site: A_p [R,K], At_p [K,R], B_p [N,R], Bt_p [R,N]."""

    def __init__(self, K: int, N: int, r_true: int, device, storage: Optional[torch.Tensor] = None, dtype=torch.bfloat16):
        self.K, self.N, self.r_true, self.R = K, N, r_true, padded_rank(r_true)
        n = self.numel(K, N, r_true)
        if storage is None:
            storage = torch.zeros(n, dtype=dtype, device=device)
        assert storage.numel() == n and storage.dtype in (torch.bfloat16, torch.float16)
        self.dtype = storage.dtype
        R = self.R
        self.storage = storage
        o = 0
        self.A_p = storage[o:o + R * K].view(R, K); o += R * K
        self.At_p = storage[o:o + K * R].view(K, R); o += K * R
        self.B_p = storage[o:o + N * R].view(N, R); o += N * R
        self.Bt_p = storage[o:o + R * N].view(R, N)
        self.versions = (-1, -1)

    @staticmethod
    def numel(K: int, N: int, r_true: int) -> int:
        return 2 * padded_rank(r_true) * (K + N)

    def site(self, lora_A: torch.Tensor, lora_B: torch.Tensor) -> PackSite:
        return PackSite(lora_A.data_ptr(), lora_B.data_ptr(), self.A_p.data_ptr(), self.At_p.data_ptr(),
                        self.B_p.data_ptr(), self.Bt_p.data_ptr(), self.K, self.N, self.r_true, self.R)


class _LoRAProjection(torch.autograd.Function):
    """y[M,N] = x[M,K] W^T + b + s (x A^T) B^T through the C ABI; W and b are frozen (model.py:137)."""

    @staticmethod
    def forward(ctx, x2, lora_A, lora_B, mod, res2=None):
        """``res2`` (optional, [M,N], the dtype of x2): the residual stream the output is added to -- in the epilogue of the same
        launch (``sdt_lora_linear_fwd_res``); its gradient is the output's."""
        lib = _lib.load()
        M, K = x2.shape
        N = mod.out_features
        code = _lib.dtype_code(x2.dtype)
        st = _lib.stream_ptr()
        y = torch.empty(M, N, dtype=x2.dtype, device=x2.device)
        ev0 = None
        if code in (_lib.SDT_BF16, _lib.SDT_F16):
            ops = mod._packed_operands(x2.dtype)
            w = mod._weight_lp(x2.dtype)
            t_save = torch.empty(M, ops.R, dtype=x2.dtype, device=x2.device)
            if PROFILE is not None:
                ev0 = _ev()
            if res2 is None:
                _lib.check(lib.sdt_lora_linear_fwd(x2.data_ptr(), w.data_ptr(), _lib.ptr(mod._bias_f32()), ops.A_p.data_ptr(),
                                                   ops.B_p.data_ptr(), mod.scaling, y.data_ptr(), t_save.data_ptr(), M, K, N,
                                                   ops.R, code, st), "sdt_lora_linear_fwd")
            else:
                _lib.check(lib.sdt_lora_linear_fwd_res(x2.data_ptr(), w.data_ptr(), _lib.ptr(mod._bias_f32()), ops.A_p.data_ptr(),
                                                       ops.B_p.data_ptr(), mod.scaling, res2.data_ptr(), y.data_ptr(),
                                                       t_save.data_ptr(), M, K, N, ops.R, code, st), "sdt_lora_linear_fwd_res")
        else:
            if res2 is not None:
                raise SdtError("the residual epilogue is bf16 / fp16 only")
            if mod.weight.dtype != torch.float32:
                raise SdtError("fp32 activations need fp32 frozen weights")
            r = mod.r
            t_save = torch.empty(M, r, dtype=torch.float32, device=x2.device)
            _lib.check(lib.sdt_lora_linear_fwd(x2.data_ptr(), mod.weight.data_ptr(), _lib.ptr(mod._bias_f32()),
                                               lora_A.data_ptr(), lora_B.data_ptr(), mod.scaling, y.data_ptr(),
                                               t_save.data_ptr(), M, K, N, r, code, st), "sdt_lora_linear_fwd")
        if ev0 is not None:
            PROFILE.append(("fwd+res" if res2 is not None else "fwd", M, K, N, ops.R, 1, ev0, _ev()))
        ctx.mod = mod
        ctx.code = code
        ctx.need_dx = x2.requires_grad
        ctx.need_w = lora_A.requires_grad or lora_B.requires_grad
        ctx.save_for_backward(x2, t_save, lora_A, lora_B)
        ctx.res_grad = res2 is not None and res2.requires_grad
        return y

    @staticmethod
    def backward(ctx, dy):
        x2, t_save, lora_A, lora_B = ctx.saved_tensors
        dx, dA, dB = _site_backward(ctx.mod, ctx.code, x2, t_save, lora_A, lora_B, dy, ctx.need_dx)
        return dx, dA, dB, None, (dy if ctx.res_grad else None)


# ---- deferred weight-gradient reductions --------------------------------------------------------------------------------------
class _WgradQueue:
    """dA / dB reductions waiting to go out as ONE ``sdt_lora_wgrad_batch`` launch.  Inside ``deferred_wgrad()`` the backward
    of a bf16 / fp16 site computes dX and G = s dY B at once (the gradient chain needs dX) but only *queues* its two token
    reductions; a full queue -- about a transformer block's sites -- or the end of the context flushes them.  Nothing reads
    dA / dB before the optimizer, so the only effect is fewer, larger launches.  The queue keeps every operand alive."""

    def __init__(self, max_sites: int):
        self.max_sites = max_sites
        self.items: list = []

    def add(self, x2, g_ws, dA, dy, t_save, dB, M, K, N, R, r_true, code) -> None:
        if self.items and (self.items[0][9:] != (R, r_true, code)):
            self.flush()                                # one launch = one padded rank / dtype
        self.items.append((x2, g_ws, dA, dy, t_save, dB, M, K, N, R, r_true, code))
        if len(self.items) >= self.max_sites:
            self.flush()

    def flush(self) -> None:
        if not self.items:
            return
        items, self.items = self.items, []
        R, r_true, code = items[0][9:]
        arr = (_lib.WgradSite * len(items))(*[
            _lib.WgradSite(x.data_ptr(), g.data_ptr(), dA.data_ptr(), dy.data_ptr(), ts.data_ptr(), dB.data_ptr(), M, K, N)
            for x, g, dA, dy, ts, dB, M, K, N, *_ in items])
        ev0 = _ev() if PROFILE is not None else None
        _lib.check(_lib.load().sdt_lora_wgrad_batch(ctypes.addressof(arr), len(items), R, r_true, code, _lib.wgrad_workspace(),
                                                    _lib.stream_ptr()), "sdt_lora_wgrad_batch")
        if ev0 is not None:
            PROFILE.append(("wgrad", 0, 0, 0, R, len(items), [(M, K, N) for *_, M, K, N, _R, _rt, _c in items], ev0, _ev()))


_wgrad_queue: Optional[_WgradQueue] = None


class deferred_wgrad:
    """``with deferred_wgrad(): loss.backward()`` -- batch the LoRA weight-gradient reductions of the backward pass (see
    ``_WgradQueue``).  Gradients are complete when the context exits.  Not re-entrant, one backward at a time."""

    def __init__(self, max_sites: Optional[int] = None):
        self.max_sites = max_sites

    def __enter__(self):
        global _wgrad_queue
        if _wgrad_queue is not None:
            raise SdtError("deferred_wgrad() is not re-entrant")
        n = self.max_sites or int(_lib.load().sdt_lora_wgrad_max_sites())
        _wgrad_queue = _WgradQueue(min(n, int(_lib.load().sdt_lora_wgrad_max_sites())))
        return _wgrad_queue

    def __exit__(self, exc_type, exc, tb):
        global _wgrad_queue
        q, _wgrad_queue = _wgrad_queue, None
        if exc_type is None:
            q.flush()
        return False


def _site_backward(mod, code, x2, t_save, lora_A, lora_B, dy, need_dx):
    """dX (optional), dA, dB of one site through ``sdt_lora_linear_bwd``.  Returns (dx, dA, dB); dA / dB are None when the
    kernels accumulated straight into the flat gradient arena."""
    lib = _lib.load()
    M, K = x2.shape
    N = mod.out_features
    st = _lib.stream_ptr()
    dy = dy.contiguous()
    if dy.dtype != x2.dtype:
        dy = dy.to(x2.dtype)
    dx = torch.empty_like(x2) if need_dx else None
    # gradient destinations: the flat arena when one is attached (no per-site copies), else fresh zeros
    direct = mod._grad_A is not None
    if direct:
        dA, dB = mod._grad_A, mod._grad_B
    else:
        dA = torch.zeros(mod.r, K, dtype=torch.float32, device=x2.device)
        dB = torch.zeros(N, mod.r, dtype=torch.float32, device=x2.device)
    if code in (_lib.SDT_BF16, _lib.SDT_F16):
        ops = mod._packed_operands(x2.dtype)
        wt = mod._weight_t_lp(x2.dtype) if need_dx else None
        g_ws = torch.empty(M, ops.R, dtype=x2.dtype, device=x2.device)
        ev0 = _ev() if PROFILE is not None else None
        q = _wgrad_queue if direct else None        # only arena gradients may be written after this backward has returned
        _lib.check(lib.sdt_lora_linear_bwd(dy.data_ptr(), x2.data_ptr(), _lib.ptr(wt), ops.At_p.data_ptr(),
                                           ops.Bt_p.data_ptr(), t_save.data_ptr(), mod.scaling, _lib.ptr(dx),
                                           g_ws.data_ptr(), None if q is not None else dA.data_ptr(),
                                           None if q is not None else dB.data_ptr(), M, K, N, ops.R, mod.r,
                                           code, _lib.wgrad_workspace(), st), "sdt_lora_linear_bwd")
        if ev0 is not None:
            PROFILE.append(("bwd", M, K, N, ops.R, 1, need_dx, ev0, _ev()))
        if q is not None:
            q.add(x2, g_ws, dA, dy, t_save, dB, M, K, N, ops.R, mod.r, code)
    else:
        g_ws = torch.empty(M, mod.r, dtype=torch.float32, device=x2.device)
        _lib.check(lib.sdt_lora_linear_bwd(dy.data_ptr(), x2.data_ptr(), mod.weight.data_ptr(), lora_A.data_ptr(),
                                           lora_B.data_ptr(), t_save.data_ptr(), mod.scaling, _lib.ptr(dx),
                                           g_ws.data_ptr(), dA.data_ptr(), dB.data_ptr(), M, K, N, mod.r, mod.r,
                                           code, None, st), "sdt_lora_linear_bwd")
    if direct:
        return dx, None, None
    return dx, dA, dB


class _LoRAGegluProjection(torch.autograd.Function):
    """``h, gate = proj(x).chunk(2, -1); h * gelu(gate)`` with ``proj`` a LoRA site (diffusers ``GEGLU`` under
    ``configs/optim_targets/lora.yaml:23-27``) as ONE launch: the activation is formed in the GEMM's epilogue
    (``sdt_lora_linear_geglu_fwd``).  ``proj`` is still written -- the backward needs it -- and the backward is the GEGLU
    gradient kernel followed by the site's usual dX / dA / dB."""

    @staticmethod
    def forward(ctx, x2, lora_A, lora_B, mod):
        lib = _lib.load()
        M, K = x2.shape
        N = mod.out_features
        I = N // 2
        code = _lib.dtype_code(x2.dtype)
        ops = mod._packed_operands(x2.dtype)
        proj = torch.empty(M, N, dtype=x2.dtype, device=x2.device)
        act = torch.empty(M, I, dtype=x2.dtype, device=x2.device)
        t_save = torch.empty(M, ops.R, dtype=x2.dtype, device=x2.device)
        ev0 = _ev() if PROFILE is not None else None
        _lib.check(lib.sdt_lora_linear_geglu_fwd(x2.data_ptr(), mod._weight_lp(x2.dtype).data_ptr(), _lib.ptr(mod._bias_f32()),
                                                 ops.A_p.data_ptr(), ops.B_p.data_ptr(), mod.scaling, proj.data_ptr(), act.data_ptr(),
                                                 t_save.data_ptr(), M, K, I, ops.R, code, _lib.stream_ptr()),
                   "sdt_lora_linear_geglu_fwd")
        if ev0 is not None:
            PROFILE.append(("fwd", M, K, N, ops.R, 1, ev0, _ev()))
        ctx.mod, ctx.code = mod, code
        ctx.need_dx = x2.requires_grad
        ctx.save_for_backward(x2, t_save, proj, lora_A, lora_B)
        return act

    @staticmethod
    def backward(ctx, dact):
        x2, t_save, proj, lora_A, lora_B = ctx.saved_tensors
        M, N = proj.shape
        d2 = dact.contiguous()
        if d2.dtype != proj.dtype:
            d2 = d2.to(proj.dtype)
        dproj = torch.empty_like(proj)
        _lib.check(_lib.load().sdt_geglu(proj.data_ptr(), d2.data_ptr(), dproj.data_ptr(), M, N // 2, 1, ctx.code, _lib.stream_ptr()),
                   "sdt_geglu")
        dx, dA, dB = _site_backward(ctx.mod, ctx.code, x2, t_save, lora_A, lora_B, dproj, ctx.need_dx)
        return dx, dA, dB, None


def geglu_projection_supported(mod, x: torch.Tensor) -> bool:
    """True when ``h * gelu(gate)`` of ``mod(x)`` can come out of the projection's own epilogue."""
    if not isinstance(mod, LoRALinear) or not x.is_cuda or x.numel() == 0 or mod.out_features % 2 != 0:
        return False
    # Opt-in (SDT_FUSED_GEGLU=1).  Measured on B200 (profiles/r02_geglu_epilogue_ab.txt): the erf GELU is ~25 instructions per
    # element and only the 8 epilogue warps of a CTA can evaluate it, so at K = 320 the epilogue becomes issue-bound (146 us against
    # 60 us projection + 45 us vectorised GEGLU pass) and at K >= 640 the 128-wide tiles it needs cost more than the pass saves.
    # The default is therefore the 224-wide projection followed by the GEGLU kernel; `geglu_projection` itself ignores the switch.
    if os.environ.get("SDT_FUSED_GEGLU", "0") != "1":
        return False
    if mod.training and mod.lora_dropout_p > 0.0:
        return False
    dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else x.dtype
    if dt != torch.bfloat16:                   # the GEGLU gradient kernel is bf16 / fp32
        return False
    M = x.numel() // x.shape[-1]
    return bool(_lib.load().sdt_lora_linear_geglu_supported(M, mod.in_features, mod.out_features // 2, padded_rank(mod.r)))


def geglu_projection(mod, x: torch.Tensor) -> torch.Tensor:
    """``h * gelu(gate)`` for ``[h | gate] = mod(x)`` -- one fused launch (``geglu_projection_supported`` says when)."""
    _lib.require_cuda(x, mod.weight, mod.lora_A)
    _lib.device_check()
    lead = x.shape[:-1]
    x2 = x.reshape(-1, mod.in_features)
    if x2.dtype != torch.bfloat16:
        x2 = x2.to(torch.bfloat16)
    act = _LoRAGegluProjection.apply(x2.contiguous(), mod.lora_A, mod.lora_B, mod)
    return act.view(*lead, mod.out_features // 2)


# ---- dropout on the rank path (loralib: ``(dropout(x) @ A.T @ B.T) * scaling``) -------------------------------------------------
_dropout_seed: dict = {}
_dropout_sites = 0


def dropout_seed(device) -> torch.Tensor:
    """The device-resident int64 seed of the LoRA dropout masks on ``device`` (one per device).  Kernels read it from memory, so a
    captured CUDA graph draws new masks on every replay as long as ``advance_dropout_seed`` runs between replays."""
    key = torch.device(device).index
    t = _dropout_seed.get(key)
    if t is None:
        t = torch.zeros(1, dtype=torch.int64, device=device)
        t.random_()                                           # from torch's default CUDA generator: follows torch.manual_seed
        _dropout_seed[key] = t
    return t


def advance_dropout_seed(device, generator=None) -> None:
    """New masks for the next step (called once per optimisation step by the trainer, outside any captured graph)."""
    dropout_seed(device).random_(generator=generator)


class _LoRADropoutProjection(torch.autograd.Function):
    """Training-mode projection of a site with ``lora_dropout > 0``: the fused kernels on the concatenated contraction
    ``[x | dropout(x)]`` (see ``csrc/dropout.cu``).  Twice the GEMM work of the plain path -- no shipped optim_target uses
    dropout, this exists so that a reference config that sets it runs instead of raising."""

    @staticmethod
    def forward(ctx, x2, lora_A, lora_B, mod):
        lib = _lib.load()
        M, K = x2.shape
        N = mod.out_features
        code = _lib.dtype_code(x2.dtype)
        st = _lib.stream_ptr()
        dev = x2.device
        mod._drop_calls += 1
        salt = (mod._drop_site << 32) | (mod._drop_calls & 0xFFFFFFFF)
        seed = dropout_seed(dev).clone()                      # the value this forward used, for the backward
        xcat = torch.empty(M, 2 * K, dtype=x2.dtype, device=dev)
        _lib.check(lib.sdt_lora_dropout(x2.data_ptr(), xcat.data_ptr(), M, K, mod.lora_dropout_p, seed.data_ptr(), salt, 0, code, st),
                   "sdt_lora_dropout")
        ops = mod._packed_operands(x2.dtype)
        a_cat = torch.zeros(ops.R, 2 * K, dtype=x2.dtype, device=dev)
        a_cat[:, K:] = ops.A_p                                # A' = [0 | A]: the rank path sees only the dropped-out half
        y = torch.empty(M, N, dtype=x2.dtype, device=dev)
        t_save = torch.empty(M, ops.R, dtype=x2.dtype, device=dev)
        _lib.check(lib.sdt_lora_linear_fwd(xcat.data_ptr(), mod._weight_cat_lp(x2.dtype).data_ptr(), _lib.ptr(mod._bias_f32()),
                                           a_cat.data_ptr(), ops.B_p.data_ptr(), mod.scaling, y.data_ptr(), t_save.data_ptr(),
                                           M, 2 * K, N, ops.R, code, st), "sdt_lora_linear_fwd")
        ctx.mod, ctx.code, ctx.salt = mod, code, salt
        ctx.need_dx = x2.requires_grad
        ctx.save_for_backward(xcat, t_save, seed)
        return y

    @staticmethod
    def backward(ctx, dy):
        xcat, t_save, seed = ctx.saved_tensors
        mod, code = ctx.mod, ctx.code
        lib = _lib.load()
        st = _lib.stream_ptr()
        M, K2 = xcat.shape
        K, N = K2 // 2, mod.out_features
        dev = xcat.device
        dy = dy.contiguous()
        if dy.dtype != xcat.dtype:
            dy = dy.to(xcat.dtype)
        ops = mod._packed_operands(xcat.dtype)
        at_cat = torch.zeros(K2, ops.R, dtype=xcat.dtype, device=dev)
        at_cat[K:] = ops.At_p
        dxcat = torch.empty(M, K2, dtype=xcat.dtype, device=dev) if ctx.need_dx else None
        g_ws = torch.empty(M, ops.R, dtype=xcat.dtype, device=dev)
        dA_cat = torch.zeros(mod.r, K2, dtype=torch.float32, device=dev)
        direct = mod._grad_A is not None
        dB = mod._grad_B if direct else torch.zeros(N, mod.r, dtype=torch.float32, device=dev)
        wt = mod._weight_cat_t_lp(xcat.dtype) if ctx.need_dx else None
        _lib.check(lib.sdt_lora_linear_bwd(dy.data_ptr(), xcat.data_ptr(), _lib.ptr(wt), at_cat.data_ptr(), ops.Bt_p.data_ptr(),
                                           t_save.data_ptr(), mod.scaling, _lib.ptr(dxcat), g_ws.data_ptr(), dA_cat.data_ptr(),
                                           dB.data_ptr(), M, K2, N, ops.R, mod.r, code, _lib.wgrad_workspace(), st),
                   "sdt_lora_linear_bwd")
        dx = None
        if ctx.need_dx:
            dx = torch.empty(M, K, dtype=xcat.dtype, device=dev)
            _lib.check(lib.sdt_lora_dropout(dxcat.data_ptr(), dx.data_ptr(), M, K, mod.lora_dropout_p, seed.data_ptr(), ctx.salt, 1,
                                            code, st), "sdt_lora_dropout")
        dA = dA_cat[:, K:]                                     # the gradient of A' = [0 | A] restricted to A
        if direct:
            mod._grad_A.add_(dA)
            return dx, None, None, None
        return dx, dA.contiguous(), dB, None


class _LoRAProjectionGroup(torch.autograd.Function):
    """G same-shape projections of ONE input (to_q / to_k / to_v on the normalised hidden states; to_k / to_v of one or two
    cross-attentions on the text context) as the work items of one ``sdt_lora_linear_fwd_group`` launch."""

    @staticmethod
    def forward(ctx, x2, mods, *lora_params):
        lib = _lib.load()
        M, K = x2.shape
        N = mods[0].out_features
        G = len(mods)
        st = _lib.stream_ptr()
        ops = [m._packed_operands(x2.dtype) for m in mods]
        R = ops[0].R
        code = _lib.dtype_code(x2.dtype)
        ys = [torch.empty(M, N, dtype=x2.dtype, device=x2.device) for _ in mods]
        ts = [torch.empty(M, R, dtype=x2.dtype, device=x2.device) for _ in mods]
        probs = (_lib.LoraProblem * G)(*[
            _lib.LoraProblem(x2.data_ptr(), m._weight_lp(x2.dtype).data_ptr(), _lib.ptr(m._bias_f32()), o.A_p.data_ptr(),
                             o.B_p.data_ptr(), y.data_ptr(), t.data_ptr())
            for m, o, y, t in zip(mods, ops, ys, ts)])
        ev0 = _ev() if PROFILE is not None else None
        _lib.check(lib.sdt_lora_linear_fwd_group(ctypes.addressof(probs), G, mods[0].scaling, M, K, N, R, code, st),
                   "sdt_lora_linear_fwd_group")
        if ev0 is not None:
            PROFILE.append(("fwd", M, K, N, R, G, ev0, _ev()))
        ctx.mods = mods
        ctx.need_dx = x2.requires_grad
        ctx.save_for_backward(x2, *ts, *lora_params)
        return tuple(ys)

    @staticmethod
    def backward(ctx, *dys):
        mods = ctx.mods
        G = len(mods)
        saved = ctx.saved_tensors
        x2, ts, lora_params = saved[0], saved[1:1 + G], saved[1 + G:]
        M, K = x2.shape
        N = mods[0].out_features
        lib = _lib.load()
        ops = [m._packed_operands(x2.dtype) for m in mods]
        R = ops[0].R
        code = _lib.dtype_code(x2.dtype)
        # every projection received a gradient and the shape qualifies: dX of all of them in ONE launch (summed sources);
        # without dX (text context) the rank projections G = s dY B of the group are the work items of one launch
        if (all(dy is not None for dy in dys)
                and lib.sdt_lora_linear_bwd_group_supported(G, int(ctx.need_dx), M, K, N, R)):
            st = _lib.stream_ptr()
            dys = [dy.contiguous() if dy.dtype == x2.dtype else dy.to(x2.dtype).contiguous() for dy in dys]
            dx = torch.empty_like(x2) if ctx.need_dx else None
            gws = [torch.empty(M, R, dtype=x2.dtype, device=x2.device) for _ in mods]
            grads, dAs, dBs = [], [], []
            for m in mods:
                if m._grad_A is not None:
                    dAs.append(m._grad_A); dBs.append(m._grad_B)
                    grads += [None, None]
                else:
                    dA = torch.zeros(m.r, K, dtype=torch.float32, device=x2.device)
                    dB = torch.zeros(N, m.r, dtype=torch.float32, device=x2.device)
                    dAs.append(dA); dBs.append(dB)
                    grads += [dA, dB]
            q = _wgrad_queue if all(m._grad_A is not None for m in mods) else None
            probs = (_lib.LoraBwdProblem * G)(*[
                _lib.LoraBwdProblem(dy.data_ptr(), x2.data_ptr(), m._weight_t_lp(x2.dtype).data_ptr() if ctx.need_dx else None,
                                    o.At_p.data_ptr(), o.Bt_p.data_ptr(), t.data_ptr(), g.data_ptr(),
                                    None if q is not None else dA.data_ptr(), None if q is not None else dB.data_ptr())
                for dy, m, o, t, g, dA, dB in zip(dys, mods, ops, ts, gws, dAs, dBs)])
            ev0 = _ev() if PROFILE is not None else None
            _lib.check(lib.sdt_lora_linear_bwd_group(ctypes.addressof(probs), G, mods[0].scaling, _lib.ptr(dx), M, K, N, R,
                                                     mods[0].r, code, _lib.wgrad_workspace(), st),
                       "sdt_lora_linear_bwd_group")
            if ev0 is not None:
                PROFILE.append(("bwd", M, K, N, R, G, ctx.need_dx, ev0, _ev()))
            if q is not None:
                for dy, m, t, g, dA, dB in zip(dys, mods, ts, gws, dAs, dBs):
                    q.add(x2, g, dA, dy, t, dB, M, K, N, R, m.r, code)
            return (dx, None, *grads)
        dx = None
        grads = []
        for g, (m, dy) in enumerate(zip(mods, dys)):
            if dy is None:                    # this output was not used downstream
                grads += [None, None]
                continue
            dx_g, dA, dB = _site_backward(m, code, x2, ts[g], lora_params[2 * g], lora_params[2 * g + 1], dy, ctx.need_dx)
            grads += [dA, dB]
            if dx_g is not None:
                dx = dx_g if dx is None else dx.add_(dx_g)
        return (dx, None, *grads)


class _LoRAProjectionMulti(torch.autograd.Function):
    """Projections of ONE input with DIFFERENT output widths as the work items of one launch (``sdt_lora_linear_fwd_multi``):
    to_k / to_v of every cross-attention of the UNet on the text context.  The input needs no gradient (frozen text
    embeddings); the backward runs the rank projections G = s dY B of same-width sites as grouped launches and queues the dA / dB
    reductions like every other site."""

    @staticmethod
    def forward(ctx, x2, mods, *lora_params):
        lib = _lib.load()
        M, K = x2.shape
        n = len(mods)
        code = _lib.dtype_code(x2.dtype)
        ops = [m._packed_operands(x2.dtype) for m in mods]
        R = ops[0].R
        ys = [torch.empty(M, m.out_features, dtype=x2.dtype, device=x2.device) for m in mods]
        ts = [torch.empty(M, R, dtype=x2.dtype, device=x2.device) for _ in mods]
        probs = (_lib.LoraProblem * n)(*[
            _lib.LoraProblem(x2.data_ptr(), m._weight_lp(x2.dtype).data_ptr(), _lib.ptr(m._bias_f32()), o.A_p.data_ptr(),
                             o.B_p.data_ptr(), y.data_ptr(), t.data_ptr())
            for m, o, y, t in zip(mods, ops, ys, ts)])
        widths = (ctypes.c_int64 * n)(*[m.out_features for m in mods])
        ev0 = _ev() if PROFILE is not None else None
        _lib.check(lib.sdt_lora_linear_fwd_multi(ctypes.addressof(probs), ctypes.addressof(widths), n, mods[0].scaling, M, K, R, code,
                                                 _lib.stream_ptr()), "sdt_lora_linear_fwd_multi")
        if ev0 is not None:
            PROFILE.append(("fwd_multi", M, K, [m.out_features for m in mods], R, n, ev0, _ev()))
        ctx.mods = mods
        ctx.save_for_backward(x2, *ts, *lora_params)
        return tuple(ys)

    @staticmethod
    def backward(ctx, *dys):
        mods = ctx.mods
        n = len(mods)
        saved = ctx.saved_tensors
        x2, ts, lora_params = saved[0], saved[1:1 + n], saved[1 + n:]
        code = _lib.dtype_code(x2.dtype)
        lib = _lib.load()
        M, K = x2.shape
        R = mods[0]._packed_operands(x2.dtype).R
        grads: list = [None] * (2 * n)
        # same-width sites whose outputs all received a gradient go out MAX_GROUP at a time: their rank projections G = s dY B are
        # the work items of one launch (sdt_lora_linear_bwd_group without dX); anything else per site
        by_width: dict = {}
        for g, (m, dy) in enumerate(zip(mods, dys)):
            if dy is not None:
                by_width.setdefault(m.out_features, []).append(g)
        for N, idx in by_width.items():
            for i in range(0, len(idx), _lib.MAX_GROUP):
                chunk = idx[i:i + _lib.MAX_GROUP]
                G = len(chunk)
                direct = all(mods[g]._grad_A is not None for g in chunk)
                if G < 2 or not direct or not lib.sdt_lora_linear_bwd_group_supported(G, 0, M, K, N, R):
                    for g in chunk:
                        _dx, dA, dB = _site_backward(mods[g], code, x2, ts[g], lora_params[2 * g], lora_params[2 * g + 1], dys[g], False)
                        grads[2 * g], grads[2 * g + 1] = dA, dB
                    continue
                st = _lib.stream_ptr()
                dcs = [dys[g].contiguous() if dys[g].dtype == x2.dtype else dys[g].to(x2.dtype).contiguous() for g in chunk]
                gws = [torch.empty(M, R, dtype=x2.dtype, device=x2.device) for _ in chunk]
                q = _wgrad_queue
                probs = (_lib.LoraBwdProblem * G)(*[
                    _lib.LoraBwdProblem(dc.data_ptr(), x2.data_ptr(), None, mods[g]._packed_operands(x2.dtype).At_p.data_ptr(),
                                        mods[g]._packed_operands(x2.dtype).Bt_p.data_ptr(), ts[g].data_ptr(), gw.data_ptr(),
                                        None if q is not None else mods[g]._grad_A.data_ptr(),
                                        None if q is not None else mods[g]._grad_B.data_ptr())
                    for g, dc, gw in zip(chunk, dcs, gws)])
                ev0 = _ev() if PROFILE is not None else None
                _lib.check(lib.sdt_lora_linear_bwd_group(ctypes.addressof(probs), G, mods[chunk[0]].scaling, None, M, K, N, R,
                                                         mods[chunk[0]].r, code, _lib.wgrad_workspace(), st), "sdt_lora_linear_bwd_group")
                if ev0 is not None:
                    PROFILE.append(("bwd", M, K, N, R, G, False, ev0, _ev()))
                if q is not None:
                    for g, dc, gw in zip(chunk, dcs, gws):
                        q.add(x2, gw, mods[g]._grad_A, dc, ts[g], mods[g]._grad_B, M, K, N, R, mods[g].r, code)
        return (None, None, *grads)


def multi_projectable(mods, x2: torch.Tensor) -> bool:
    """True when ``mods`` (LoRA Linear sites of one in-width, rank, scaling and bias-ness, any out-widths) can share ONE launch on
    ``x2`` [M,K] that needs no gradient."""
    if not (1 <= len(mods) <= _lib.MAX_MULTI) or not all(isinstance(m, LoRALinear) for m in mods):
        return False
    if not x2.is_cuda or x2.requires_grad or os.environ.get("SDT_MULTI_CONTEXT", "1") == "0":
        return False
    dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else x2.dtype
    if dt not in (torch.bfloat16, torch.float16):
        return False
    m0 = mods[0]
    if not all(m.in_features == m0.in_features and m.r == m0.r and m.scaling == m0.scaling and (m.bias is None) == (m0.bias is None)
               and not (m.training and m.lora_dropout_p > 0.0) for m in mods):
        return False
    widths = (ctypes.c_int64 * len(mods))(*[m.out_features for m in mods])
    return bool(_lib.load().sdt_lora_linear_fwd_multi_supported(len(mods), x2.shape[0], x2.shape[1], ctypes.addressof(widths),
                                                                padded_rank(m0.r)))


def project_multi(mods, x: torch.Tensor):
    """``[m(x) for m in mods]`` for sites of different out-widths on one gradient-free input -- one launch
    (``multi_projectable`` says when)."""
    mods = list(mods)
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    _lib.require_cuda(x2, *(m.weight for m in mods))
    _lib.device_check()
    dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else x2.dtype
    if x2.dtype != dt:
        x2 = x2.to(dt)
    params = [p for m in mods for p in (m.lora_A, m.lora_B)]
    ys = _LoRAProjectionMulti.apply(x2.contiguous(), mods, *params)
    return [y.view(*lead, m.out_features) for y, m in zip(ys, mods)]


def groupable(mods, x2: torch.Tensor) -> bool:
    """True when ``mods`` can share one grouped launch on ``x2`` [M,K]: LoRA sites of one shape, rank, scaling and bias-ness,
    bf16 tensor-core path.  Anything else goes through the per-site launches (same kernels, one problem each)."""
    if not (2 <= len(mods) <= _lib.MAX_GROUP) or not all(isinstance(m, LoRALinear) for m in mods):
        return False
    m0 = mods[0]
    if not x2.is_cuda or x2.shape[0] == 0:
        return False
    dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else x2.dtype
    if dt not in (torch.bfloat16, torch.float16):
        return False
    return all(m.in_features == m0.in_features and m.out_features == m0.out_features and m.r == m0.r
               and m.scaling == m0.scaling and (m.bias is None) == (m0.bias is None) and m.lora_dropout_p == 0.0
               for m in mods)


def project_group(mods, x: torch.Tensor):
    """``[m(x) for m in mods]`` -- one launch when the sites are ``groupable``."""
    mods = list(mods)
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    if not groupable(mods, x2):
        return [m(x) for m in mods]
    _lib.require_cuda(x2, *(m.weight for m in mods))
    _lib.device_check()
    dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else x2.dtype
    if x2.dtype != dt:
        x2 = x2.to(dt)
    params = [p for m in mods for p in (m.lora_A, m.lora_B)]
    ys = _LoRAProjectionGroup.apply(x2.contiguous(), mods, *params)
    return [y.view(*lead, m.out_features) for y, m in zip(ys, mods)]


class _LoRABase(nn.Module):
    """State shared by the Linear and 1x1-Conv2d LoRA modules."""

    def _init_lora(self, in_features: int, out_features: int, rank: int, alpha, dropout: float):
        if rank <= 0:
            raise SdtError("LoRA rank must be positive")
        self.in_features, self.out_features, self.r = in_features, out_features, rank
        self.scaling = alpha / rank          # python float fixed at construction (loralib 0.1)
        self.lora_dropout_p = float(dropout)
        self.lora_A = nn.Parameter(torch.zeros(rank, in_features))
        self.lora_B = nn.Parameter(torch.zeros(out_features, rank))
        nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B)
        # runtime caches (not state): bf16 copies of the frozen weight and the packed LoRA operands
        self._w_cache = None
        self._wt_cache = None
        self._b_cache = None
        self._ops: Optional[PackedOperands] = None
        self._ops_external = False           # True when a LoraArena owns (and refreshes) the packed operands
        self._arena_ref = None               # weakref to that arena (refresh_packed_operands)
        self._grad_A = None
        self._grad_B = None
        global _dropout_sites
        _dropout_sites += 1
        self._drop_site = _dropout_sites      # salt of this site's dropout masks
        self._drop_calls = 0
        self._wcat_cache = None
        self._wcat_t_cache = None

    # ---- frozen operand caches ------------------------------------------------------------------
    def _weight_2d(self) -> torch.Tensor:
        return self.weight.view(self.out_features, self.in_features)

    def _weight_lp(self, dtype=torch.bfloat16) -> torch.Tensor:
        """the frozen weight [N,K] in the 16-bit compute format (bf16 or fp16); cached"""
        w = self.weight
        key = (w.data_ptr(), w._version, dtype)
        if self._w_cache is None or self._w_cache[0] != key:
            self._w_cache = (key, self._weight_2d().detach().to(dtype).contiguous())
        return self._w_cache[1]

    def _weight_t_lp(self, dtype=torch.bfloat16) -> torch.Tensor:
        """its transpose [K,N] (the B-type operand of the input-gradient GEMM); cached"""
        w = self.weight
        key = (w.data_ptr(), w._version, dtype)
        if self._wt_cache is None or self._wt_cache[0] != key:
            self._wt_cache = (key, self._weight_2d().detach().to(dtype).t().contiguous())
        return self._wt_cache[1]

    _weight_bf16, _weight_t_bf16 = _weight_lp, _weight_t_lp     # older call sites / tools

    def _weight_cat_lp(self, dtype) -> torch.Tensor:
        """[W | 0] [N, 2K]: the base weight on the concatenated contraction of the dropout path; cached"""
        w = self.weight
        key = (w.data_ptr(), w._version, dtype)
        if self._wcat_cache is None or self._wcat_cache[0] != key:
            wl = self._weight_lp(dtype)
            self._wcat_cache = (key, torch.cat([wl, torch.zeros_like(wl)], dim=1).contiguous())
        return self._wcat_cache[1]

    def _weight_cat_t_lp(self, dtype) -> torch.Tensor:
        """its transpose [2K, N] = [W^T ; 0]; cached"""
        w = self.weight
        key = (w.data_ptr(), w._version, dtype)
        if self._wcat_t_cache is None or self._wcat_t_cache[0] != key:
            wt = self._weight_t_lp(dtype)
            self._wcat_t_cache = (key, torch.cat([wt, torch.zeros_like(wt)], dim=0).contiguous())
        return self._wcat_t_cache[1]

    def _bias_f32(self) -> Optional[torch.Tensor]:
        b = self.bias
        if b is None:
            return None
        if b.dtype == torch.float32:
            return b.detach()
        key = (b.data_ptr(), b._version)
        if self._b_cache is None or self._b_cache[0] != key:
            self._b_cache = (key, b.detach().float())
        return self._b_cache[1]

    def _packed_operands(self, dtype=torch.bfloat16) -> PackedOperands:
        if self._ops_external:
            if self._ops.dtype != dtype:      # first launch in the other 16-bit format: the arena re-packs every site once
                arena = self._arena_ref() if self._arena_ref is not None else None
                if arena is None:
                    raise SdtError("packed LoRA operands belong to an arena that no longer exists")
                arena.set_compute_dtype(dtype)
            return self._ops       # refreshed by LoraArena.pack() after every optimizer step
        A, B = self.lora_A, self.lora_B
        if self._ops is None or self._ops.storage.device != A.device or self._ops.dtype != dtype:
            self._ops = PackedOperands(self.in_features, self.out_features, self.r, A.device, dtype=dtype)
        ver = (A._version, B._version)
        if self._ops.versions != ver:
            _pack_sites([self._ops.site(A.detach(), B.detach())], self._ops.R * (self.in_features + self.out_features), A.device,
                        dtype)
            self._ops.versions = ver
        return self._ops

    def _project(self, x2: torch.Tensor, res2: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``res2`` (optional [M, out]): added to the projection -- in the launch's epilogue on the 16-bit path, with torch otherwise."""
        _lib.require_cuda(x2, self.weight, self.lora_A, res2)
        _lib.device_check()
        if torch.is_autocast_enabled():
            x2 = x2.to(torch.get_autocast_dtype("cuda"))
        if x2.shape[0] == 0:                 # empty batch: nothing to launch (F.linear returns an empty tensor too)
            y = x2.new_zeros(0, self.out_features) + 0.0 * (self.lora_A.sum() + self.lora_B.sum()).to(x2.dtype)
            return y if res2 is None else y + res2
        if self.training and self.lora_dropout_p > 0.0:       # loralib: dropout acts on the rank path's input, training mode only
            if x2.dtype == torch.float32:
                raise SdtError("lora dropout > 0 runs on the bf16 / fp16 tensor-core path only (fp32 is the parity path)")
            y = _LoRADropoutProjection.apply(x2.contiguous(), self.lora_A, self.lora_B, self)
            return y if res2 is None else y + res2
        if res2 is not None and (x2.dtype == torch.float32 or res2.dtype != x2.dtype or not res2.is_contiguous()):
            return _LoRAProjection.apply(x2.contiguous(), self.lora_A, self.lora_B, self) + res2
        return _LoRAProjection.apply(x2.contiguous(), self.lora_A, self.lora_B, self, res2)

    def extra_repr(self) -> str:
        return f"in={self.in_features}, out={self.out_features}, r={self.r}, scaling={self.scaling}"


class LoRALinear(_LoRABase):
    """Replacement for ``loralib.Linear`` as adapted by ``get_lora`` (``modules/lora.py:13-14``)."""

    def __init__(self, in_features: int, out_features: int, rank=4, alpha=1, dropout=0.0):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_features, in_features), requires_grad=False)
        self.bias = None
        self._init_lora(in_features, out_features, rank, alpha, dropout)

    def forward(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``residual`` (optional, the shape of the output): returns ``proj(x) + residual`` with the add done in the projection's
        epilogue (the block's residual connection)."""
        lead = x.shape[:-1]
        res2 = None if residual is None else residual.reshape(-1, self.out_features)
        y = self._project(x.reshape(-1, self.in_features), res2)
        return y.view(*lead, self.out_features)


class LoRAConv2d(_LoRABase):
    """Replacement for ``loralib.Conv2d`` with ``kernel_size == 1`` (``modules/lora.py:15-16``), the shape of
    SD1.x ``proj_in`` / ``proj_out``.  loralib materialises ``W + s (B A).view(W.shape)`` and convolves with default
    stride / padding; for a 1x1 kernel that is the same projection applied per pixel, so the channels-last token
    matrix goes through the same fused kernel and the result is returned as an NCHW tensor in channels-last memory.
    """

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, rank=4, alpha=1, dropout=0.0):
        super().__init__()
        if kernel_size != 1:
            raise SdtError(f"LoRA Conv2d with kernel_size={kernel_size} is not implemented (SD proj_in/proj_out are 1x1)")
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, 1, 1), requires_grad=False)
        self.bias = None
        self._init_lora(in_channels, out_channels, rank, alpha, dropout)

    def forward(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        b, c, h, w = x.shape
        tokens = x.permute(0, 2, 3, 1).reshape(b * h * w, c)        # a view when x is channels-last
        res2 = None
        if residual is not None:                                    # [b, out, h, w]; a token view when it is channels-last
            res2 = residual.permute(0, 2, 3, 1).reshape(b * h * w, self.out_channels)
        y = self._project(tokens, res2)
        return y.view(b, h, w, self.out_channels).permute(0, 3, 1, 2)


def get_lora(module: nn.Linear | nn.Conv2d, rank=4, alpha=1, dropout=0.):
    """``modules/lora.py:12-27`` with the CUDA-backed modules in place of loralib's."""
    if isinstance(module, nn.Linear):
        lora = LoRALinear(module.in_features, module.out_features, rank, alpha, dropout)
    elif isinstance(module, nn.Conv2d):
        lora = LoRAConv2d(module.in_channels, module.out_channels, module.kernel_size[0], rank, alpha, dropout)
    else:
        raise Exception("Unexpected module type")

    lora.weight = module.weight
    lora.bias = module.bias
    lora.lora_A.requires_grad = True
    lora.lora_B.requires_grad = True
    lora.register_buffer("lora_alpha", torch.tensor(alpha, dtype=torch.int32))

    return lora.to(module.weight.device)


def get_linears(module: nn.Module):
    """``modules/lora.py:6-9`` (unused by the reference; kept for API completeness)."""
    for name, sub in module.named_children():
        if isinstance(sub, nn.Linear):
            yield name, sub


def refresh_packed_operands(module: nn.Module) -> None:
    """Bring the bf16 tensor-core operands of every LoRA site under ``module`` back in line with the fp32 masters.

    The kernels read packed bf16 copies (``A_p / At_p / B_p / Bt_p``), refreshed after ``optimizer.step`` by
    ``LoraArena.pack()``.  Anything ELSE that writes the masters -- ``ExponentialMovingAverage.apply`` /
    ``average_parameters``, ``load_state_dict``, manual ``p.data.copy_`` -- must call this, otherwise the next bf16 forward
    silently runs with the old factors (``.data`` writes do not bump ``Parameter._version``)."""
    arenas = {}
    for _, m in lora_modules(module):
        if m._ops_external:
            arena = m._arena_ref() if m._arena_ref is not None else None
            if arena is not None:
                arenas[id(arena)] = arena
        elif m._ops is not None:
            m._ops.versions = (-1, -1)         # repacked lazily by the next forward
    for arena in arenas.values():
        arena.pack()


def lora_modules(module: nn.Module):
    for name, sub in module.named_modules():
        if isinstance(sub, _LoRABase):
            yield name, sub
