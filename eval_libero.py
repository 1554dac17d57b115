#!/usr/bin/env python
"""Entry point mirroring reference eval_libero.py:29-121 for the action-inference path.

The LIBERO simulator (MuJoCo) is out of scope; `--synthetic_rollout_steps N` drives `ModelWrapper.step` with synthetic
observations and reports p50 / p99 per-action latency (the latency metric of BASELINE.json).  A real env can call
`dreamvla_b200.utils.eval_utils_calvin.ModelWrapper.step` exactly where the reference calls `model.step` (:264).
"""
from __future__ import annotations

import json
import time

import torch

from dreamvla_b200.models import DreamVLA
from dreamvla_b200.utils.arguments_utils import get_parser, model_kwargs
from dreamvla_b200.utils.eval_utils_calvin import ModelWrapper


def main(args):
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    model = DreamVLA(finetune_type=args.finetune_type, clip_device="cpu", vit_checkpoint_path=args.vit_checkpoint_path,
                     **model_kwargs(args)).bfloat16().to(dev)
    model._init_model_type()
    if args.resume_from_checkpoint:
        ck = torch.load(args.resume_from_checkpoint, map_location="cpu")["model_state_dict"]
        model.load_state_dict({k[len("module."):] if k.startswith("module.") else k: v for k, v in ck.items()}, strict=False)
    model.eval()
    wrapper = ModelWrapper(model, history_len=args.sequence_length, action_pred_steps=args.action_pred_steps, device=dev)
    n = args.synthetic_rollout_steps or 100
    g = torch.Generator().manual_seed(args.seed)
    text = torch.zeros(77, dtype=torch.long)
    text[0], text[1:6], text[6] = 49406, torch.randint(1, 49406, (5,), generator=g), 49407
    lat = []
    for i in range(n):
        img, grip, obs = torch.randn(3, 224, 224, generator=g), torch.randn(3, 224, 224, generator=g), torch.randn(15, generator=g)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        wrapper.step(img, grip, obs, text)
        torch.cuda.synchronize()
        lat.append((time.perf_counter() - t0) * 1e3)
    lat_s = sorted(lat[min(5, n // 10):])
    print(json.dumps({"metric": "action_inference_latency_ms", "p50": lat_s[len(lat_s) // 2], "p99": lat_s[int(len(lat_s) * 0.99) - 1],
                      "steps": n, "seq_len": args.sequence_length}))


if __name__ == "__main__":
    main(get_parser().parse_args())
