"""Which torch (non-libdvla) kernels run inside one eager train step, and from which aten op / shapes: torch.profiler table
sorted by CUDA time.  python tools/torch_op_profile.py [--batch 8]"""
import argparse
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from dreamvla_b200.utils.train_utils import StepConfig, TrainStep, synthetic_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
a = ap.parse_args()
dev = torch.device("cuda", 0)
cfg = bench.CONFIGS["calvin"]
scfg = StepConfig(**cfg["step"])
model = bench.build_model(cfg, dev, 0.1)
step = TrainStep(model, scfg)
heads = dict(cfg["heads"], flow_mask=scfg.flow_as_mask)
batch = synthetic_batch(scfg, a.batch, dev, seed=1, heads=heads)
for _ in range(2):
    step(batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    step(batch)
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=40,
                                                           max_shapes_column_width=70))

# where do the torch copy / elementwise kernels come from?  (aten ops with CUDA time, with input shapes and Python stack)
rows = []
for e in prof.events():
    if e.name in ("aten::copy_", "aten::add", "aten::cat", "aten::add_", "aten::contiguous", "aten::clone", "aten::to") \
            and e.device_time_total > 8:
        rows.append((e.device_time_total, e.name, str(e.input_shapes)[:90], " <- ".join(f.split("/")[-1] for f in (e.stack or [])[:5])))
from collections import defaultdict
agg = defaultdict(lambda: [0.0, 0])
for t, n, sh, st in rows:
    agg[(n, sh, st)][0] += t
    agg[(n, sh, st)][1] += 1
for (n, sh, st), (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
    print(f"{t/1e3:8.3f} ms n={c:4d} {n:16s} {sh:90s} {st}")
