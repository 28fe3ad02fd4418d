"""Per-site A/B on this GPU: ONE fused launch of this repo against the sequence the reference executes for the same site
(``modules/lora.py:14`` -> loralib 0.1 under autocast: ``F.linear`` + ``x @ A.T`` + ``@ B.T`` + ``* scaling`` + ``+`` = 5
launches forward, and autograd's chain of them backward: dY W, dY B, (dY B) A, the adds, dA, dB).

Forward and backward, the twelve (M, K, N) of BASELINE cfg2, rank 16 and 64.  CUDA-event time over the whole
sequence (L2-warm, 20 iterations after warm-up), so launch gaps between the torch kernels count -- they are what a training
step pays.  Prints a table; `python tools/site_vs_torch.py > profiles/r02_site_vs_torch.txt`.
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scal_sdt_b200 import get_lora  # noqa: E402

dev = torch.device("cuda:0")
SHAPES = [(32768, 320, 320), (32768, 320, 2560), (32768, 1280, 320), (8192, 640, 640), (8192, 640, 5120), (8192, 2560, 640),
          (2048, 1280, 1280), (2048, 1280, 10240), (2048, 5120, 1280), (512, 1280, 1280), (512, 1280, 10240), (512, 5120, 1280)]


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3          # microseconds


def graphed(fn):
    """the same sequence replayed from a CUDA graph: no host launch gaps (the best case for the torch sequence)"""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g.replay


def main():
    print(f"{'M':>6} {'K':>5} {'N':>6} {'r':>3} | {'fused fwd':>9} {'torch fwd':>9} {'x':>5} | {'fused f+b':>9} {'torch f+b':>9} {'x':>5} | "
          f"{'fused TF/s':>10} {'torch TF/s':>10}   (us per site, CUDA graph replay on both sides; TF/s over fwd+bwd)")
    for r in (16, 64):
        for M, K, N in SHAPES:
            torch.manual_seed(0)
            base = torch.nn.Linear(K, N, bias=True).to(dev).to(torch.bfloat16).requires_grad_(False)
            ours = get_lora(base, r, r)
            with torch.no_grad():
                ours.lora_B.normal_(0, 0.1)
            A32, B32 = ours.lora_A.detach().clone().requires_grad_(True), ours.lora_B.detach().clone().requires_grad_(True)
            w, b, s = base.weight, base.bias, ours.scaling
            x = torch.randn(M, K, device=dev, dtype=torch.bfloat16, requires_grad=True)
            dy = torch.randn(M, N, device=dev, dtype=torch.bfloat16)

            def torch_fwd():
                # what autocast makes of loralib's forward: every matmul in bf16 (the fp32 masters are cast per call)
                res = F.linear(x, w, b)
                return res + (x @ A32.to(torch.bfloat16).T @ B32.to(torch.bfloat16).T) * s

            def torch_fb():
                A32.grad = B32.grad = x.grad = None
                torch_fwd().backward(dy)

            def ours_fwd():
                return ours(x)

            def ours_fb():
                x.grad = None
                ours(x).backward(dy)

            with torch.no_grad():
                t_of = timed(graphed(lambda: ours(x.detach())))
                t_tf = timed(graphed(lambda: F.linear(x.detach(), w, b) + (x.detach() @ A32.detach().to(torch.bfloat16).T
                                                                           @ B32.detach().to(torch.bfloat16).T) * s))
            t_ob = timed(graphed(ours_fb))
            t_tb = timed(graphed(torch_fb))
            fl = 4.0 * M * K * N + 6.0 * M * r * (K + N)
            print(f"{M:>6} {K:>5} {N:>6} {r:>3} | {t_of:9.1f} {t_tf:9.1f} {t_tf / t_of:5.2f} | {t_ob:9.1f} {t_tb:9.1f} {t_tb / t_ob:5.2f} | "
                  f"{fl / t_ob / 1e6:10.0f} {fl / t_tb / 1e6:10.0f}")
            del base, ours, x, dy


if __name__ == "__main__":
    main()
