"""(Residual add +) LayerNorm kernels in isolation on the token shapes of cfg2, buffers rotated so that every launch reads HBM.
With a second library path the same measurement runs on that build too (A/B of two builds in one process):

    python tools/ln_bench.py [other/libsdt_b200.so]
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scal_sdt_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
SHAPES = [(32768, 320), (8192, 640), (2048, 1280), (512, 1280)]
HBM = 6535.7
VP, F, I64, I = ctypes.c_void_p, ctypes.c_float, ctypes.c_int64, ctypes.c_int


def bind(path):
    lib = ctypes.CDLL(path)
    lib.sdt_layer_norm_fwd.restype = I
    lib.sdt_layer_norm_fwd.argtypes = [VP, VP, VP, VP, VP, VP, VP, I64, I, F, VP]
    lib.sdt_layer_norm_bwd.restype = I
    lib.sdt_layer_norm_bwd.argtypes = [VP, VP, VP, VP, VP, VP, I64, I, VP]
    return lib


def run(lib, name, other=None):
    st = torch.cuda.current_stream().cuda_stream
    for M, C in SHAPES:
        copies = max(2, int(400e6 // (M * C * 2 * 4)) + 1)
        x = [torch.randn(M, C, device=dev).bfloat16() for _ in range(copies)]
        r = [torch.randn(M, C, device=dev).bfloat16() for _ in range(copies)]
        xs = [torch.empty(M, C, device=dev, dtype=torch.bfloat16) for _ in range(copies)]
        y = [torch.empty(M, C, device=dev, dtype=torch.bfloat16) for _ in range(copies)]
        g = torch.randn(C, device=dev) * 0.2 + 1
        b = torch.randn(C, device=dev) * 0.2
        stats = torch.empty(M, 2, device=dev)
        for res in (True, False):
            for backward in (False, True):
                def one(j, L=lib):
                    c = j % copies
                    if backward:
                        rc = L.sdt_layer_norm_bwd(xs[c].data_ptr(), x[c].data_ptr(), r[c].data_ptr() if res else None, g.data_ptr(), stats.data_ptr(),
                                                  y[c].data_ptr(), M, C, st)
                    else:
                        rc = L.sdt_layer_norm_fwd(x[c].data_ptr(), r[c].data_ptr() if res else None, g.data_ptr(), b.data_ptr(),
                                                  xs[c].data_ptr() if res else None, y[c].data_ptr(), stats.data_ptr(), M, C, 1e-5, st)
                    assert rc == 0
                for j in range(copies + 1):           # the forward also fills xs / stats for the backward
                    one(j)
                a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                iters = 20
                torch.cuda.synchronize()
                a.record()
                for j in range(iters):
                    one(j)
                e.record()
                torch.cuda.synchronize()
                us = a.elapsed_time(e) * 1e3 / iters
                nbytes = 2.0 * M * C * ((4 if res else 2) if not backward else (4 if res else 3))
                line = f"{name:5s} M={M:6d} C={C:5d} {'bwd' if backward else 'fwd'} {'+res' if res else '    '}: {us:6.1f} us {nbytes / us / 1e3:6.0f} GB/s = {nbytes / us / 1e3 / HBM:.2f} of HBM copy"
                if other is not None:
                    one(0)
                    mine = y[0].clone()
                    one(0, other)
                    torch.cuda.synchronize()
                    ne = (mine != y[0]).sum().item()
                    line += f" | vs the other build: {ne} of {mine.numel()} elements differ, max |diff| {(mine.float() - y[0].float()).abs().max().item():.3g}"
                print(line)


new = bind(str(_lib.library_path()))
oth = bind(sys.argv[1]) if len(sys.argv) > 1 else None
for rnd in range(2 if oth is not None else 1):
    if oth is not None:
        run(oth, "other")
    run(new, "new", oth if rnd == 0 else None)
