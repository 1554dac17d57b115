"""Deterministic synthetic weights / inputs shared by the golden generator, the oracle tests and the GPU parity tests.

Every tensor depends only on (key name, shape, seed) through a per-key torch CPU generator, so the reference (here), the
oracle and the CUDA model (on the GPU box) reconstruct identical values without shipping weights.
"""
from __future__ import annotations

import zlib

import torch

BASE = dict(sequence_length=2, num_resampler_query=16, num_obs_token_per_image=9, action_pred_steps=3,
            transformer_layers=2, hidden_dim=1024, transformer_heads=16, phase="finetune", track_label_patch_size=8,
            attn_implementation="sdpa", batch=1, weight_seed=11, input_seed=12, draw_seed=13)

CASES = {
    # CALVIN-style: all five world-knowledge heads + DiT head (BASELINE config C2, shrunk in depth / window)
    "calvin_allheads": dict(BASE, obs_pred=True, depth_pred=True, trajectory_pred=True, dino_feat_pred=True,
                            sam_feat_pred=True, use_dit_head=True),
    # LIBERO-style: world heads off, DiT head (C3)
    "libero_dit": dict(BASE, obs_pred=False, depth_pred=False, trajectory_pred=False, dino_feat_pred=False,
                       sam_feat_pred=False, use_dit_head=True, sequence_length=3),
    # pretrain-style: obs head only, MLP action head, goal-conditioned mask variants (C5)
    "pretrain_mlp": dict(BASE, obs_pred=True, depth_pred=False, trajectory_pred=False, dino_feat_pred=False,
                         sam_feat_pred=False, use_dit_head=False, sequence_length=5, atten_goal=4, atten_goal_state=True,
                         atten_only_obs=True, attn_robot_proprio_state=True, mask_l_obs_ratio=0.0, phase="evaluate"),
}

# Cases pinned on the CPU only (reference -> golden -> oracle): constructor switches of the shipped scripts that the three
# GPU cases above do not exercise.  Not in CASES so that the GPU suite keeps its size.
CPU_ONLY_CASES = {
    # LIBERO scripts pass --gripper_width: 8-dim state, the 2-dim gripper opening goes straight into gripper_state_encoder
    "libero_gripper_width": dict(BASE, obs_pred=False, depth_pred=False, trajectory_pred=False, dino_feat_pred=False,
                                 sam_feat_pred=False, use_dit_head=True, sequence_length=3, gripper_width=True),
}

# Flow-matching action head (`--use_fm`, eval_libero.py:75; ActionModelFM / FMDiffusion): reference -> golden -> oracle on the
# CPU, CUDA path in tests/test_model_gpu.py::test_flow_matching_head
FM_CASES = {
    "libero_fm": dict(BASE, obs_pred=False, depth_pred=False, trajectory_pred=False, dino_feat_pred=False,
                      sam_feat_pred=False, use_dit_head=True, sequence_length=3, use_fm=True, weight_seed=41, input_seed=42,
                      draw_seed=43),
}

# Full-depth cases: BASELINE.json's C2 (CALVIN, 24 layers, S=10, five world heads + DiT, L=1290) and C3 (LIBERO, S=7,
# --gripper_width, world heads off, L=273) at batch 1.  GPU parity only (tests/test_full_depth_gpu.py); the reference runs
# them on the CPU in fp32 in ~1 min each when the goldens are generated.
FULL_CASES = {
    "calvin_full": dict(BASE, obs_pred=True, depth_pred=True, trajectory_pred=True, dino_feat_pred=True, sam_feat_pred=True,
                        use_dit_head=True, transformer_layers=24, sequence_length=10, weight_seed=21, input_seed=22,
                        draw_seed=23),
    "libero_full": dict(BASE, obs_pred=False, depth_pred=False, trajectory_pred=False, dino_feat_pred=False,
                        sam_feat_pred=False, use_dit_head=True, transformer_layers=24, sequence_length=7, gripper_width=True,
                        weight_seed=31, input_seed=32, draw_seed=33),
}

CTOR_KEYS = ("sequence_length", "num_resampler_query", "num_obs_token_per_image", "obs_pred", "atten_only_obs",
             "attn_robot_proprio_state", "atten_goal", "atten_goal_state", "mask_l_obs_ratio", "action_pred_steps",
             "transformer_layers", "hidden_dim", "transformer_heads", "phase", "gripper_width", "pred_num", "depth_pred",
             "trajectory_pred", "track_label_patch_size", "dino_feat_pred", "sam_feat_pred", "use_dit_head",
             "attn_implementation", "use_fm")


def ctor_kwargs(cfg):
    return {k: cfg[k] for k in CTOR_KEYS if k in cfg}


def _gen(key: str, seed: int) -> torch.Generator:
    return torch.Generator(device="cpu").manual_seed((zlib.crc32(key.encode()) * 2654435761 + seed * 97) % (2 ** 63))


KEEP = ("attention_mask", "position_embedding", "pos_embed", "logit_scale")


def synth_tensor(key: str, template: torch.Tensor, seed: int) -> torch.Tensor:
    if any(k in key for k in KEEP) and "transformer_backbone_position_embedding" not in key \
            and "positional_embedding" not in key:
        return template.detach().clone().float()
    if not template.is_floating_point():
        return template.detach().clone()
    g = _gen(key, seed)
    shape = tuple(template.shape)
    if template.dim() <= 1:
        if key.endswith("weight"):          # LayerNorm gains
            return 1.0 + 0.1 * torch.randn(shape, generator=g)
        return 0.05 * torch.randn(shape, generator=g)
    if "token" in key and "embedding" not in key and "projector" not in key:   # learned query / mask / cls tokens
        return 0.5 * torch.randn(shape, generator=g)
    if "latents" in key or "uncondition" in key:
        return 0.5 * torch.randn(shape, generator=g)
    if "transformer_backbone_position_embedding" in key or "positional_embedding" in key:
        return 0.1 * torch.randn(shape, generator=g)
    if "token_embedding" in key:
        return 0.05 * torch.randn(shape, generator=g)
    fan = 1
    for s in shape:
        fan *= s
    fan_in_out = shape[0] + fan // shape[0]
    std = (2.0 / fan_in_out) ** 0.5
    return std * torch.randn(shape, generator=g)


def synth_state_dict(template_sd, seed: int):
    return {k: synth_tensor(k, v, seed) for k, v in template_sd.items()}


def synth_inputs(cfg):
    B, S = cfg["batch"], cfg["sequence_length"]
    g = torch.Generator(device="cpu").manual_seed(cfg["input_seed"])
    image_primary = torch.randn(B, S, 3, 224, 224, generator=g)
    image_wrist = torch.randn(B, S, 3, 224, 224, generator=g)
    if cfg.get("gripper_width", False):       # 6 arm dims + 2 finger positions (dreamvla_model.py:656-664)
        state = torch.randn(B, S, 8, generator=g) * 0.5
    else:
        state = torch.randn(B, S, 7, generator=g) * 0.5
        state[..., 6] = (torch.rand(B, S, generator=g) < 0.5).float()
    text = torch.zeros(B, 77, dtype=torch.long)
    for b in range(B):
        k = int(torch.randint(3, 21, (1,), generator=g))
        ids = torch.randint(1, 49406, (k,), generator=g)
        text[b, 0] = 49406
        text[b, 1:1 + k] = ids
        text[b, 1 + k] = 49407
    text_token = text.unsqueeze(1).repeat(1, S, 1)
    action_label = torch.rand(B, S, cfg["action_pred_steps"], 7, generator=g) * 2 - 1
    action_label[..., 6] = (action_label[..., 6] > 0).float()
    return dict(image_primary=image_primary, image_wrist=image_wrist, state=state, text_token=text_token,
                action_label=action_label)


def synth_labels(cfg, window=None):
    """Label tensors with the collator's output contract, seeded."""
    B, S = cfg["batch"], cfg["sequence_length"]
    W = window or (S + 3)
    g = torch.Generator(device="cpu").manual_seed(cfg["input_seed"] + 1000)
    return dict(
        images_primary=torch.randn(B, W, 3, 224, 224, generator=g), images_wrist=torch.randn(B, W, 3, 224, 224, generator=g),
        depth_primary=torch.rand(B, W, 1, 224, 224, generator=g) * 4.9 + 0.1,
        depth_wrist=torch.rand(B, W, 1, 224, 224, generator=g) * 4.9 + 0.1,
        dino_primary=torch.randn(B, W, 256, 768, generator=g), dino_wrist=torch.randn(B, W, 256, 768, generator=g),
        sam_primary=torch.randn(B, W, 256, 256, generator=g), sam_wrist=torch.randn(B, W, 256, 256, generator=g),
        tracks=torch.randn(B, W, 784, 2, generator=g) * 2, tracks_gripper=torch.randn(B, W, 784, 2, generator=g) * 2)


def subsample(t: torch.Tensor, n: int = 4096) -> torch.Tensor:
    f = t.detach().float().flatten()
    if f.numel() <= n:
        return f.clone()
    idx = torch.linspace(0, f.numel() - 1, n).long()
    return f[idx].clone()
