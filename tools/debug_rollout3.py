"""Value-level comparison of the incremental rollout variants against the full-window wrapper, step by step."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dreamvla_b200.utils.eval_utils_calvin import ModelWrapper  # noqa: E402
from tests import synth  # noqa: E402
from tests.test_rollout_gpu import build, observations  # noqa: E402

dev = torch.device("cuda:0")
cfg = synth.CASES[os.environ.get("DBG_CASE", "libero_dit")]
S = cfg["sequence_length"]
model = build(cfg, dev)
text, obs = observations(S + 3, seed=11)
full = ModelWrapper(model, history_len=S, device=dev, use_cuda_graph=False)
variants = {"eager_full_tokens": dict(use_cuda_graph=False, prune=False), "graph_full_tokens": dict(use_cuda_graph=True, prune=False),
            "eager_pruned": dict(use_cuda_graph=False, prune=True), "graph_pruned": dict(use_cuda_graph=True, prune=True)}
wr = {k: ModelWrapper(model, history_len=S, device=dev, incremental=True, **kw) for k, kw in variants.items()}
g = torch.Generator().manual_seed(4)
for i in range(len(obs)):
    noise = torch.randn(S, 3, 7, generator=g)
    a = full.step(*obs[i], text, sample_noise=noise).astype(np.float32)
    line = f"step {i} full {a[:3]}"
    for k, w in wr.items():
        b = w.step(*obs[i], text, sample_noise=noise).astype(np.float32)
        tk = torch.stack(list(w.tok_queue))
        line += f" | {k}: maxdiff {np.nanmax(np.abs(a - b)):.3e} nan={int(np.isnan(b).sum())} tok_nan={int(torch.isnan(tk.float()).sum())} tok_absmax={float(tk.float().abs().max()):.2f}"
    print(line, flush=True)
