"""Device timeline of ONE eager data-parallel train step on rank 0 (torch.profiler / CUPTI; nsys is not in this image):
where the NCCL all-reduce kernels sit relative to the backward kernels, how long each runs, how much of the exchange is
exposed after the last backward kernel, and how much slower this package's kernels run while an all-reduce is in flight.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/ddp_timeline.py [--batch 8]"""
import argparse
import os
import sys

import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from dreamvla_b200.utils.distributed_utils import configure_nccl  # noqa: E402
from dreamvla_b200.utils.train_utils import StepConfig, TrainStep, synthetic_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    configure_nccl()
    dist.init_process_group("nccl", device_id=dev)
cfg = bench.CONFIGS["calvin"]
scfg = StepConfig(**cfg["step"])
model = bench.build_model(cfg, dev, 0.1)
step = TrainStep(model, scfg, world_size=world)
heads = dict(cfg["heads"], flow_mask=scfg.flow_as_mask)
batch = synthetic_batch(scfg, a.batch, dev, seed=1 + rank, heads=heads)
for _ in range(3):
    step(batch)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
    torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(batch)
    torch.cuda.synchronize()
if rank == 0:
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    ev.sort(key=lambda e: e.time_range.start)
    t0 = ev[0].time_range.start
    nccl = [e for e in ev if "nccl" in e.name.lower()]
    comp = [e for e in ev if "nccl" not in e.name.lower() and "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
    opt = [e for e in comp if "adamw" in e.name or "sumsq" in e.name or "grad_clip" in e.name]
    bwd_end = max(e.time_range.end for e in comp if e not in opt and e.time_range.start < (opt[0].time_range.start if opt else 1e30))
    span = ev[-1].time_range.end - t0
    print(f"[timeline] world={world} per-GPU batch={a.batch}: step span {span/1e3:.2f} ms, {len(comp)} compute kernels "
          f"({sum(e.time_range.end - e.time_range.start for e in comp)/1e3:.2f} ms busy), {len(nccl)} NCCL kernels "
          f"({sum(e.time_range.end - e.time_range.start for e in nccl)/1e3:.2f} ms busy)")
    print(f"[timeline] last forward/backward kernel ends at {(bwd_end - t0)/1e3:.2f} ms; optimizer kernels start at "
          f"{((opt[0].time_range.start - t0)/1e3 if opt else float('nan')):.2f} ms -> exposed exchange "
          f"{((opt[0].time_range.start - bwd_end)/1e3 if opt else float('nan')):.2f} ms")
    for e in nccl:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        inside = [c for c in comp if c.time_range.start < e.time_range.end and c.time_range.end > e.time_range.start]
        print(f"[timeline]   NCCL {e.name[:60]:60s} start {s/1e3:8.2f} ms  dur {d/1e3:7.2f} ms  overlapping compute kernels {len(inside)}")
    # slowdown of this package's kernels while an all-reduce is in flight: same kernel name, mean duration inside / outside
    import collections
    inside_t, outside_t = collections.defaultdict(list), collections.defaultdict(list)
    for c in comp:
        hit = any(c.time_range.start < e.time_range.end and c.time_range.end > e.time_range.start for e in nccl)
        (inside_t if hit else outside_t)[c.name.split("(")[0][:70]].append(c.time_range.end - c.time_range.start)
    tot_in = sum(sum(v) for v in inside_t.values())
    print(f"[timeline] compute time inside NCCL windows {tot_in/1e3:.2f} ms; per kernel family (n inside, mean inside / mean outside):")
    for k, v in sorted(inside_t.items(), key=lambda kv: -sum(kv[1]))[:10]:
        o = outside_t.get(k)
        if o:
            print(f"[timeline]   {k:70s} n={len(v):4d}  {sum(v)/len(v):8.1f} us / {sum(o)/len(o):8.1f} us")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
