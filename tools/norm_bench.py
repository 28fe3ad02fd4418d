"""Per-shape timings of the GroupNorm / LayerNorm kernels on the shapes of the SD1.5 UNet step (batch 8, 64x64 latents)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scal_sdt_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
GN = [(8, 320, 64, 64), (8, 640, 64, 64), (8, 960, 64, 64), (8, 320, 32, 32), (8, 640, 32, 32), (8, 960, 32, 32), (8, 1280, 32, 32),
      (8, 1920, 32, 32), (8, 640, 16, 16), (8, 1280, 16, 16), (8, 1920, 16, 16), (8, 2560, 16, 16), (8, 1280, 8, 8), (8, 2560, 8, 8)]
if os.environ.get("NORM_BENCH_ONE"):
    GN = GN[:1]


def kernel_times(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
    out = {}
    for e in prof.events():
        if str(getattr(e, "device_type", "")).endswith("CUDA") and "Mem" not in e.name:
            k = e.name.split("(")[0].replace("void sdt::", "")
            out.setdefault(k, []).append(e.time_range.end - e.time_range.start)
    return {k: sum(v) / len(v) for k, v in out.items()}


st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print("GroupNorm (silu) -- us per kernel; MB = bytes of one tensor")
for (B, C, H, W) in GN:
    x = torch.randn(B, H * W, C, device=dev).bfloat16()
    d = torch.randn(B, H * W, C, device=dev).bfloat16()
    y = torch.empty_like(x)
    g = torch.ones(C, device=dev)
    b = torch.zeros(C, device=dev)
    stats = torch.empty(B, 32, 2, device=dev)
    bst = torch.empty(int(lib.sdt_group_norm_workspace_floats(B, 32)), device=dev)

    def fwd():
        _lib.check(lib.sdt_group_norm_nhwc(x.data_ptr(), None, g.data_ptr(), b.data_ptr(), stats.data_ptr(), y.data_ptr(), B, H * W, C, 32, 1e-5, 1, bst.data_ptr(), bst.numel(), st))

    def bwd():
        _lib.check(lib.sdt_group_norm_nhwc_bwd(x.data_ptr(), None, d.data_ptr(), g.data_ptr(), b.data_ptr(), stats.data_ptr(), bst.data_ptr(), bst.numel(),
                                               y.data_ptr(), B, H * W, C, 32, 1e-5, 1, st))
    tf, tb = kernel_times(fwd), kernel_times(bwd)
    mb = x.numel() * 2 / 1e6
    s = " ".join(f"{k.split('<')[0]}{'B' if 'true,' in k.split('<')[1][:6] else 'F'}={v:6.1f}" for k, v in {**tf, **tb}.items())
    ideal_f, ideal_b = 2 * mb / 6535.7 * 1e3, 3 * mb / 6535.7 * 1e3
    print(f"{B}x{C}x{H}x{W} {mb:6.1f} MB  {s}   ideal fwd {ideal_f:5.1f} bwd {ideal_b:5.1f}")
print("LayerNorm with residual")
for (M, C) in ([(32768, 320)] if os.environ.get("NORM_BENCH_ONE") else [(32768, 320), (8192, 640), (2048, 1280), (512, 1280)]):
    x = torch.randn(M, C, device=dev).bfloat16()
    r = torch.randn(M, C, device=dev).bfloat16()
    xs, y, dx = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    g = torch.ones(C, device=dev)
    b = torch.zeros(C, device=dev)
    stats = torch.empty(M, 2, device=dev)

    def fwd():
        _lib.check(lib.sdt_layer_norm_fwd(x.data_ptr(), r.data_ptr(), g.data_ptr(), b.data_ptr(), xs.data_ptr(), y.data_ptr(), stats.data_ptr(), M, C, 1e-5, st))

    def bwd():
        _lib.check(lib.sdt_layer_norm_bwd(xs.data_ptr(), y.data_ptr(), r.data_ptr(), g.data_ptr(), stats.data_ptr(), dx.data_ptr(), M, C, st))
    tf, tb = kernel_times(fwd), kernel_times(bwd)
    mb = x.numel() * 2 / 1e6
    print(f"{M}x{C} {mb:6.1f} MB fwd {list(tf.values())[0]:6.1f} us ({4 * mb / list(tf.values())[0]:5.2f} TB/s)  bwd {list(tb.values())[0]:6.1f} us ({4 * mb / list(tb.values())[0]:5.2f} TB/s)")
