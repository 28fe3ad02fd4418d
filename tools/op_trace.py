"""Which torch elementwise / copy ops does one training step still launch, on which shapes and strides?
(TorchDispatchMode over one eager step, autograd forced onto the calling thread so the backward is seen too.)"""
import collections
import os
import sys

import torch
from torch.utils._python_dispatch import TorchDispatchMode

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from scal_sdt_b200 import GradExchange  # noqa: E402

dev = torch.device("cuda:0")
tr = bench.build_trainer(bench.WORKLOAD, dev, GradExchange(0, 1))
batches = [{k: v.to(dev) for k, v in b.items()} for b in bench.synthetic_batches(bench.WORKLOAD, 2, 8, 0, False)]
for i in range(2):
    tr.step(batches[i % 2])
torch.cuda.synchronize()
SKIP = ("convolution", "scaled_dot", "_cudnn", "view", "reshape", "permute", "transpose", "detach", "empty", "alias", "expand",
        "unsqueeze", "t.default", "slice", "split", "select", "as_strided", "squeeze", "unsafe", "_to_copy")
log = collections.Counter()


def desc(a):
    if isinstance(a, torch.Tensor):
        dense = a.is_contiguous() or (a.dim() == 4 and a.is_contiguous(memory_format=torch.channels_last))
        return f"{tuple(a.shape)}{'' if dense else ' STRIDED' + str(tuple(a.stride()))}:{str(a.dtype)[6:]}"
    return None


class Trace(TorchDispatchMode):
    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        name = str(func)
        if not any(s in name for s in SKIP):
            ds = [d for d in (desc(a) for a in args) if d]
            big = any(isinstance(a, torch.Tensor) and a.numel() >= 65536 for a in args)
            if big:
                log[(name, " , ".join(ds))] += 1
        return func(*args, **(kwargs or {}))


torch.autograd.set_multithreading_enabled(False)
with Trace():
    tr.step(batches[0])
torch.cuda.synchronize()
agg = collections.Counter()
for (name, d), n in log.items():
    agg[name] += n
print("ops on tensors >= 64k elements, one step:")
for name, n in agg.most_common():
    print(f"{n:5d}  {name}")
print()
for (name, d), n in sorted(log.items(), key=lambda kv: (kv[0][0], -kv[1])):
    if any(k in name for k in ("add", "copy", "mul", "sum", "cat", "div", "sub", "fill", "zero", "clone", "contiguous")):
        print(f"{n:4d}  {name:28s} {d}")
