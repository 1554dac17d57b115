"""torchrun --nproc-per-node 2 tools/ddp_overlap_check.py
Two ranks, different data, identical init, identical RNG streams in both runs: the all-reduced flat gradient of ONE
micro-step with the backward-overlapped segment all-reduce (DVLA_AR_OVERLAP=1) must equal the one of a single all-reduce
after backward, segment by segment; then three full steps must leave identical parameters on every rank.
Run with DVLA_GEMM_SPLITK=0 so that every kernel is bit-reproducible."""
import copy
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from dreamvla_b200.utils.train_utils import StepConfig, TrainStep, synthetic_batch  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = bench.CONFIGS["calvin"]
scfg = StepConfig(**cfg["step"])
heads = dict(cfg["heads"], flow_mask=scfg.flow_as_mask)
base = bench.build_model(cfg, dev, 0.0, layers=4)          # no dropout: the two runs must see the same masks
batch = synthetic_batch(scfg, 2, dev, seed=77 + rank, heads=heads)


def ranks_agree(t):
    g = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(g, t)
    return all(torch.equal(g[0], x) for x in g)


def run(overlap, steps):
    torch.manual_seed(4321 + rank)          # the DiT head draws its noise / timesteps from the global generator
    m = copy.deepcopy(base)
    st = TrainStep(m, scfg, world_size=world)
    st.overlap = bool(overlap)
    dist.broadcast(st.flat.P, src=0)
    st.flat.lr.fill_(scfg.learning_rate)
    st.micro_step(batch)                     # forward, backward, all-reduce; no optimiser step
    torch.cuda.synchronize()
    G = st.flat.G.float().clone()
    info = (list(st.flat.seg_end), st.flat.n_big, st.flat.n)
    g_agree = ranks_agree(G)
    st.flat.G.zero_()
    losses = [float(st(batch)) for _ in range(steps)]
    torch.cuda.synchronize()
    P = st.flat.P.float().clone()
    p_agree = ranks_agree(P)
    del st, m
    return G, P, losses, info, g_agree, p_agree


G0, P0, l0, info, ga0, pa0 = run(False, 3)
G2, P2, l2, _, ga2, pa2 = run(False, 3)       # same thing again: the run-to-run noise floor (fp32 atomics in LN / bias grads)
G1, P1, l1, _, ga1, pa1 = run(True, 3)
seg_end, n_big, n = info
if rank == 0:
    print(f"segment ends: {seg_end} | last big segment ends at {n_big} | 1-D gradients up to {n}")
    bounds = [0] + list(seg_end) + [n_big, n]
    labels = [f"seg{i}" for i in range(len(seg_end) + 1)] + ["small"]
    for name, lo, hi in zip(labels, bounds[:-1], bounds[1:]):
        d = (G0[lo:hi] - G1[lo:hi]).abs()
        dn = (G0[lo:hi] - G2[lo:hi]).abs()
        print(f"  reduced gradient {name}: max|plain-overlap| = {float(d.max()):.3e} ({int((d > 0).sum())} elements differ)   "
              f"noise floor max|plain-plain| = {float(dn.max()):.3e} ({int((dn > 0).sum())})   of {hi - lo}")
    print(f"ranks agree on reduced G: plain {ga0} overlap {ga1};  on parameters after 3 steps: plain {pa0} overlap {pa1}")
    print(f"losses plain   {l0}\nlosses overlap {l1}")
    noise = float((G0 - G2).abs().max())
    ok = float((G0 - G1).abs().max()) <= 4.0 * noise + 1e-12 and ga0 and ga1 and pa0 and pa1
    print(f"max|P_plain - P_overlap| after 3 steps = {float((P0 - P1).abs().max()):.3e}   max|P_plain - P_plain'| = {float((P0 - P2).abs().max()):.3e}")
    print("DDP_OVERLAP_CHECK", "PASS" if ok else "FAIL")
base = None
from dreamvla_b200.utils.distributed_utils import shutdown_distributed  # noqa: E402
shutdown_distributed()
