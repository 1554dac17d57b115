"""torchrun --nproc-per-node 2 tools/ddp_overlap_check.py
Two ranks, different data, identical init: three train steps with the gradient all-reduce overlapped with backward
(segment hooks) must leave exactly the parameters of three steps with one all-reduce after backward.  Run with
DVLA_GEMM_SPLITK=0 so that every kernel is bit-reproducible."""
import copy
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from dreamvla_b200.utils.train_utils import GraphedTrainStep, StepConfig, TrainStep, synthetic_batch  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = bench.CONFIGS["calvin"]
scfg = StepConfig(**cfg["step"])
heads = dict(cfg["heads"], flow_mask=scfg.flow_as_mask)
base = bench.build_model(cfg, dev, 0.0, layers=4)          # no dropout: the two runs must see the same masks
batch = synthetic_batch(scfg, 2, dev, seed=77 + rank, heads=heads)


def run(overlap, graphed):
    os.environ["DVLA_AR_OVERLAP"] = "1" if overlap else "0"
    torch.manual_seed(4321 + rank)          # the DiT head draws its noise / timesteps from the global generator
    m = copy.deepcopy(base)
    st = TrainStep(m, scfg, world_size=world)
    dist.broadcast(st.flat.P, src=0)
    step = GraphedTrainStep(st, batch, warmup=1) if graphed else st
    losses = [float(step(batch)) for _ in range(3)]
    torch.cuda.synchronize()
    out = st.flat.P.float().clone(), st.flat.names, st.flat.seg_end, st.flat.n_big
    del step, st, m
    return losses, out


l0, (p0, names, seg_end, n_big) = run(False, False)
l1, (p1, _, _, _) = run(True, False)
l2, (p2, _, _, _) = run(True, True)
# graphed runs one extra (warm-up) step: compare eager overlap vs eager plain exactly, graphed vs eager loosely via losses
d = float((p0 - p1).abs().max())


def ranks_agree(p):
    g = [torch.zeros_like(p) for _ in range(world)]
    dist.all_gather(g, p)
    return all(torch.equal(g[0], t) for t in g)


consistent = ranks_agree(p1)
if rank == 0:
    print(f"ranks hold identical parameters without overlap: {ranks_agree(p0)}")
else:
    ranks_agree(p0)
if rank == 0:
    print(f"segments: seg_end={seg_end} n_big={n_big} n={p0.numel()}")
    print(f"losses plain   {l0}\nlosses overlap {l1}\nlosses overlap+graph (after 1 warm-up step) {l2}")
    print(f"max |P_plain - P_overlap| = {d:.3e}   ranks hold identical parameters: {consistent}")
    print("DDP_OVERLAP_CHECK", "PASS" if d == 0.0 and consistent else "FAIL")
base = None
from dreamvla_b200.utils.distributed_utils import shutdown_distributed  # noqa: E402
shutdown_distributed()
