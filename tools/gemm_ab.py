"""A/B timings of the fused LoRA GEMM variants (CUPTI kernel durations, L2-warm and L2-cold) on the shapes of the step."""
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
         if str(getattr(e, "device_type", "")).endswith("CUDA") and "lora_gemm" in e.name]
    names = {e.name.split("(")[0].replace("void sdt::", "") for e in prof.events() if "lora_gemm" in e.name}
    return sum(d) / len(d), ",".join(sorted(names))


def case(M, K, N, R, G, settings):
    xs = torch.randn(M, K, device=dev).bfloat16()
    ws = [(torch.randn(N, K, device=dev) / K ** 0.5).bfloat16() for _ in range(G)]
    bs = [torch.randn(N, device=dev) for _ in range(G)]
    As = [(torch.randn(R, K, device=dev) / K ** 0.5).bfloat16() for _ in range(G)]
    Bs = [(torch.randn(N, R, device=dev) * 0.1).bfloat16() for _ in range(G)]
    ys = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(G)]
    ts = [torch.empty(M, R, device=dev, dtype=torch.bfloat16) for _ in range(G)]
    probs = (_lib.LoraProblem * G)(*[_lib.LoraProblem(xs.data_ptr(), ws[g].data_ptr(), bs[g].data_ptr(), As[g].data_ptr(), Bs[g].data_ptr(),
                                                      ys[g].data_ptr(), ts[g].data_ptr()) for g in range(G)])
    import ctypes
    st = torch.cuda.current_stream().cuda_stream

    def run():
        if G == 1:
            _lib.check(lib.sdt_lora_linear_fwd(xs.data_ptr(), ws[0].data_ptr(), bs[0].data_ptr(), As[0].data_ptr(), Bs[0].data_ptr(), 0.5,
                                               ys[0].data_ptr(), ts[0].data_ptr(), M, K, N, R, 1, st))
        else:
            _lib.check(lib.sdt_lora_linear_fwd_group(ctypes.addressof(probs), G, 0.5, M, K, N, R, 1, st))
    fl = G * (2.0 * M * K * N + 2.0 * M * R * (K + N))
    out = []
    ref_out = None
    for name, kv in settings:
        for k in (11, 12, 13, 14, 20, 22, 30, 31):
            lib.sdt_debug_set(k, 0)
        for k, v in kv.items():
            lib.sdt_debug_set(k, v)
        warm, kn = kernel_us(run)
        cold, _ = kernel_us(run, cold=True)
        same = ""
        if ref_out is None:
            ref_out = [y.clone() for y in ys] + [t.clone() for t in ts]
        else:
            same = " bit-equal to the first setting" if all(torch.equal(a, b) for a, b in zip(ref_out, ys + ts)) else " DIFFERS from the first setting"
        out.append(f"{name}: warm {warm:6.1f} us {fl / warm / 1e6:6.0f} TF/s | cold {cold:6.1f} us {fl / cold / 1e6:6.0f} TF/s [{kn}]{same}")
    for k in (11, 12, 13, 14, 20, 22, 30, 31):
        lib.sdt_debug_set(k, 0)
    print(f"M={M} K={K} N={N} R={R} G={G}")
    for o in out:
        print("   ", o)


SET = [("auto  ", {}), ("single", {11: 1}), ("pair  ", {14: 64}), ("pair160", {14: 64, 12: 1})]
if len(sys.argv) > 1 and sys.argv[1] == "wide":
    SET = [("auto    ", {}), ("160-wide", {12: 1}), ("224 >=1280", {22: 1280}), ("224 >=1280 gs1", {22: 1280, 20: 1})]
if len(sys.argv) > 1 and sys.argv[1] == "ts":        # A operand through tensor memory (tcgen05.cp + TS-mode UMMAs) vs shared memory
    SET = [("SS", {}), ("TS", {30: 1}), ("SS", {}), ("TS", {30: 1})]
if len(sys.argv) > 1 and sys.argv[1] == "dt":        # double tiles (joint K loop of two column tiles, X landed once) vs one tile at a time
    SET = [("single tiles", {31: 1}), ("double, K>=1280", {31: 1280}), ("single tiles", {31: 1}), ("double, K>=1280", {31: 1280})]
    for shp in [(32768, 1280, 320, 16, 1), (8192, 2560, 640, 16, 1), (32768, 2560, 320, 16, 1), (8192, 5120, 640, 16, 1), (8192, 1280, 640, 16, 1),
                (98304, 2560, 320, 16, 1), (8192, 2560, 640, 64, 1), (32768, 2560, 320, 64, 1)]:
        case(*shp, SET)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == "groups":
    SET = [("auto", {}), ("gs1 ", {20: 1}), ("gs2 ", {20: 2}), ("gs3 ", {20: 3}), ("gs4 ", {20: 4}), ("gs6 ", {20: 6})]
SHAPES_BWD = [(32768, 320, 1280, 16, 1), (8192, 640, 2560, 16, 1), (2048, 1280, 5120, 16, 1), (32768, 2560, 320, 16, 1),
              (8192, 5120, 640, 16, 1), (2048, 10240, 1280, 16, 1)]      # dX of ff.net.2 / ff.net.0.proj: (M, contraction, outputs)
if len(sys.argv) > 2 and sys.argv[2] == "bwd":
    if sys.argv[1] in ("wide", "ts"):
        SHAPES_BWD = SHAPES_BWD + [(2048, 1280, 1280, 16, 1), (2048, 5120, 1280, 16, 1), (2048, 1280, 1280, 16, 3), (8192, 640, 5120, 16, 1)]
    for shp in SHAPES_BWD:
        case(*shp, SET)
    sys.exit(0)
for shp in [(32768, 320, 320, 16, 1), (32768, 320, 320, 16, 3), (32768, 320, 2560, 16, 1), (32768, 1280, 320, 16, 1),
            (8192, 640, 640, 16, 1), (8192, 640, 640, 16, 3), (8192, 640, 5120, 16, 1), (2048, 1280, 1280, 16, 1),
            (2048, 1280, 1280, 16, 3), (2048, 1280, 10240, 16, 1), (616, 768, 1280, 16, 4), (616, 768, 320, 16, 4)]:
    case(*shp, SET)
