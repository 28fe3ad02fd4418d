"""torch.profiler breakdown of one training step (which host-model ops dominate outside the hot path)."""
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
for i in range(3):
    tr.step(batches[i % 2])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    tr.step(batches[0])
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=45))
print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=45, max_name_column_width=50,
                                                          max_shapes_column_width=70))
