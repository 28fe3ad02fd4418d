"""Per-kernel GPU time of ONE replayed training step (CUPTI via torch.profiler): totals by kernel name, busy time vs span."""
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from scal_sdt_b200 import GradExchange  # noqa: E402

dev = torch.device("cuda:0")
wl = bench.WORKLOADS[os.environ.get("SDT_WORKLOAD", "cfg2")]
tr = bench.build_trainer(wl, dev, GradExchange(0, 1))
batches = [{k: v.to(dev) for k, v in b.items()} for b in bench.synthetic_batches(wl, 2, wl["batch"], 0, False)]
for i in range(3):
    tr.step(batches[i % 2])
torch.cuda.synchronize()
graph = "--no-graph" not in sys.argv
if graph:
    tr.enable_cuda_graph(batches[0])
    step = tr.graphed_step
else:
    step = tr.step
for i in range(3):
    step(batches[i % 2])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(batches[0])
    torch.cuda.synchronize()
evs = [e for e in prof.events() if str(getattr(e, "device_type", "")).endswith("CUDA")]
agg = collections.defaultdict(lambda: [0, 0.0])
t0 = min(e.time_range.start for e in evs)
t1 = max(e.time_range.end for e in evs)
busy = 0.0
for e in evs:
    d = e.time_range.end - e.time_range.start
    a = agg[e.name[:110]]
    a[0] += 1
    a[1] += d
    busy += d
print(f"graph={graph} events={len(evs)} span={1e-3 * (t1 - t0):.2f} ms  sum of kernel durations={1e-3 * busy:.2f} ms")
for name, (n, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{1e-3 * d:8.3f} ms {n:5d} x {d / n:8.1f} us  {name}")
