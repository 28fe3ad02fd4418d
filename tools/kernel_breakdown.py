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
evs = [e for e in prof.events() if str(getattr(e, "device_type", "")).endswith("CUDA") and e.time_range.end > e.time_range.start]
# Under programmatic dependent launch a kernel is resident while its predecessor drains, so raw durations overlap; the time a
# kernel ADDS to the step is its exclusive time on the timeline (bench.exclusive_durations)
ks = sorted(((e.name, e.time_range.start, e.time_range.end - e.time_range.start) for e in evs), key=lambda k: k[1])
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
t0 = min(k[1] for k in ks)
t1 = max(k[1] + k[2] for k in ks)
for name, excl, raw in bench.exclusive_durations(ks):
    a = agg[name[:110]]
    a[0] += 1
    a[1] += excl
    a[2] += raw
print(f"graph={graph} events={len(ks)} span={1e-3 * (t1 - t0):.2f} ms  exclusive kernel time={1e-3 * sum(a[1] for a in agg.values()):.2f} ms  "
      f"(sum of raw durations {1e-3 * sum(a[2] for a in agg.values()):.2f} ms)")
print("  exclusive ms  launches   avg excl us   (raw ms)  kernel")
for name, (n, d, raw) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{1e-3 * d:8.3f} ms {n:5d} x {d / n:8.1f} us  ({1e-3 * raw:7.3f})  {name}")
