"""A/B of the weight-gradient kernel's split heuristic (CUPTI durations, L2-warm and cold) on the shapes of the step."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scal_sdt_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def kernel_us(fn, iters=8, cold=False):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(iters):
            if cold:
                flush.zero_()
            fn()
        torch.cuda.synchronize()
    d = [e.time_range.end - e.time_range.start for e in prof.events()
         if str(getattr(e, "device_type", "")).endswith("CUDA") and "wgrad" in e.name]
    return sum(d) / len(d)


for (M, K, N, R) in [(32768, 320, 320, 16), (8192, 640, 640, 16), (2048, 1280, 1280, 16), (32768, 320, 2560, 16), (2048, 5120, 1280, 16),
                     (616, 768, 1280, 16), (512, 1280, 1280, 16)]:
    x = torch.randn(M, K, device=dev).bfloat16()
    dy = torch.randn(M, N, device=dev).bfloat16()
    wt = torch.randn(K, N, device=dev).bfloat16()
    At = torch.randn(K, R, device=dev).bfloat16()
    Bt = torch.randn(R, N, device=dev).bfloat16()
    ts = torch.randn(M, R, device=dev).bfloat16()
    g = torch.empty(M, R, device=dev, dtype=torch.bfloat16)
    dA = torch.zeros(R, K, device=dev)
    dB = torch.zeros(N, R, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    WS = 0 if os.environ.get('SDT_WGRAD_ATOMIC') == '1' else _lib.wgrad_workspace()

    def run():
        _lib.check(lib.sdt_lora_linear_bwd(dy.data_ptr(), x.data_ptr(), None, At.data_ptr(), Bt.data_ptr(), ts.data_ptr(), 0.5, None,
                                           g.data_ptr(), dA.data_ptr(), dB.data_ptr(), M, K, N, R, R, 1, WS, st))
    mb = 2.0 * M * (K + N + 2 * R) / 1e6
    res = []
    for mc, ps in [(8, 2), (16, 2), (32, 2), (8, 1), (16, 1), (4, 2), (8, 3)]:
        lib.sdt_debug_set(15, mc | (ps << 8))
        res.append(f"min{mc}/sm{ps}: {kernel_us(run):5.1f}|{kernel_us(run, cold=True):5.1f}")
    lib.sdt_debug_set(15, 0)
    print(f"M={M} K={K} N={N} ({mb:5.1f} MB, HBM floor {mb / 6535.7 * 1e3:4.1f} us)  warm|cold us  " + "  ".join(res))
