"""Which aten::add calls of a training step run the NON-vectorised elementwise kernel?  (shapes + strides per call)"""
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from scal_sdt_b200 import GradExchange  # noqa: E402

dev = torch.device("cuda:0")
tr = bench.build_trainer(bench.WORKLOAD, dev, GradExchange(0, 1))
batches = [{k: v.to(dev) for k, v in b.items()} for b in bench.synthetic_batches(bench.WORKLOAD, 2, 8, 0, False)]
for i in range(2):
    tr.step(batches[i % 2])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    tr.step(batches[0])
    torch.cuda.synchronize()
agg = collections.Counter()
tim = collections.Counter()
for e in prof.events():
    if e.name in ("aten::add", "aten::add_", "aten::copy_", "aten::mul", "aten::sum") and e.kernels:
        k = e.kernels[0].name
        kind = "VEC" if "vectorized" in k else ("ELT" if "elementwise_kernel" in k else k[:24])
        strides = getattr(e, "input_strides", None) or getattr(e, "concrete_inputs", None)
        key = (e.name, kind, str(e.input_shapes)[:90], str(strides)[:110] if kind != "VEC" else "")
        agg[key] += 1
        tim[key] += sum(kk.duration for kk in e.kernels)
for key, n in sorted(agg.items(), key=lambda kv: -tim[kv[0]]):
    print(f"{n:4d} x {tim[key] / n:6.1f} us  {key[0]:11s} {key[1]:4s} {key[2]}  {key[3]}")
