"""GEGLU kernels in isolation: GB/s of sdt_geglu (forward / backward) on the four FF shapes of cfg2, buffers rotated so that every
launch reads from HBM.  With a second library path the same measurement runs on that build too (A/B of two builds in one process):

    python tools/geglu_bench.py [other/libsdt_b200.so]
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scal_sdt_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
SHAPES = [(32768, 1280), (8192, 2560), (2048, 5120), (512, 5120)]
HBM = 6535.7


def bind(path):
    lib = ctypes.CDLL(path)
    lib.sdt_geglu.restype = ctypes.c_int
    lib.sdt_geglu.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                              ctypes.c_void_p]
    return lib


def run(lib, name):
    st = torch.cuda.current_stream().cuda_stream
    for M, I in SHAPES:
        copies = max(2, int(400e6 // (M * I * 2 * 5)) + 1)
        proj = [torch.randn(M, 2 * I, device=dev).bfloat16() for _ in range(copies)]
        dout = [torch.randn(M, I, device=dev).bfloat16() for _ in range(copies)]
        out = [torch.empty(M, I, device=dev, dtype=torch.bfloat16) for _ in range(copies)]
        dproj = [torch.empty(M, 2 * I, device=dev, dtype=torch.bfloat16) for _ in range(copies)]
        for backward, nbytes in ((0, 6.0 * M * I), (1, 10.0 * M * I)):
            def one(j):
                c = j % copies
                if backward:
                    rc = lib.sdt_geglu(proj[c].data_ptr(), dout[c].data_ptr(), dproj[c].data_ptr(), M, I, 1, _lib.SDT_BF16, st)
                else:
                    rc = lib.sdt_geglu(proj[c].data_ptr(), None, out[c].data_ptr(), M, I, 0, _lib.SDT_BF16, st)
                assert rc == 0
            for j in range(3):
                one(j)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 20
            torch.cuda.synchronize()
            a.record()
            for j in range(iters):
                one(j)
            b.record()
            torch.cuda.synchronize()
            us = a.elapsed_time(b) * 1e3 / iters
            print(f"{name:5s} M={M:6d} I={I:5d} {'bwd' if backward else 'fwd'}: {us:7.1f} us  {nbytes / us / 1e3:7.0f} GB/s = {nbytes / us / 1e3 / HBM:.2f} of HBM copy")
        if name == "new" and len(sys.argv) > 1:
            # how far the two builds are apart (bf16 outputs)
            ref = bind(sys.argv[1])
            o2, d2 = torch.empty_like(out[0]), torch.empty_like(dproj[0])
            ref.sdt_geglu(proj[0].data_ptr(), None, o2.data_ptr(), M, I, 0, _lib.SDT_BF16, st)
            ref.sdt_geglu(proj[0].data_ptr(), dout[0].data_ptr(), d2.data_ptr(), M, I, 1, _lib.SDT_BF16, st)
            lib.sdt_geglu(proj[0].data_ptr(), None, out[0].data_ptr(), M, I, 0, _lib.SDT_BF16, st)
            lib.sdt_geglu(proj[0].data_ptr(), dout[0].data_ptr(), dproj[0].data_ptr(), M, I, 1, _lib.SDT_BF16, st)
            torch.cuda.synchronize()
            def diff(a, b):
                ne = (a != b).sum().item()
                return f"{ne} of {a.numel()} elements differ, max |diff| {(a.float() - b.float()).abs().max().item():.3g}"
            print(f"      vs the other build: fwd {diff(o2, out[0])}; bwd {diff(d2, dproj[0])}")


if len(sys.argv) > 1:
    run(bind(sys.argv[1]), "other")
run(bind(str(_lib.library_path())), "new")
if len(sys.argv) > 1:
    run(bind(sys.argv[1]), "other")
    run(bind(str(_lib.library_path())), "new")
