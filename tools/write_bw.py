"""How fast can this GPU WRITE?  (fill / copy of tensors of the sizes the GEMM epilogues produce)"""
import torch
dev = torch.device("cuda:0")
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3
for mb in (21, 84, 168, 672, 2048):
    n = mb * (1 << 20) // 2
    y = torch.empty(n, dtype=torch.bfloat16, device=dev)
    x = torch.empty(n, dtype=torch.bfloat16, device=dev)
    tf = t(lambda: y.fill_(1.0))
    tc = t(lambda: y.copy_(x))
    print(f"{mb:5d} MiB  fill {tf*1e6:7.1f} us = {n*2/tf/1e12:5.2f} TB/s written   copy {tc*1e6:7.1f} us = {2*n*2/tc/1e12:5.2f} TB/s (r+w)")
