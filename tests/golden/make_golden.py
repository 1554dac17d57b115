"""Generates the golden vectors under tests/golden/ FROM THE UNMODIFIED REFERENCE (/root/reference + oracle/ref_shims.py).

Run in the build container (the reference is not on the GPU box):   python tests/golden/make_golden.py

For each case: build the reference `DreamVLA` (fp32, CPU), overwrite its state_dict with the deterministic synthetic
state of `synth.synth_state_dict` (so the fixture needs no weights, only a seed), run the reference forward on the
deterministic synthetic inputs of `synth.synth_inputs`, and store a fixed subsample of every output
(`synth.subsample`) plus the scalar losses computed by the reference's own loss code path semantics.

Sampled tensors inside the reference forward (diffusion noise / timesteps / label-drop ids, DDIM start noise) are drawn
by torch in the reference; they are captured by wrapping torch.randn_like / randint / rand / randn for the duration of
the call and stored in the fixture, so the oracle and the CUDA path consume exactly the same draws.
The ViT patch-token permutation (vit_mae.py:157-182 at ratio 0) is forced to identity for the main vectors and,
separately, left random to record the (fp32-noise-level) invariance in `perm_invariance_rel`.
"""
from __future__ import annotations

import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
from tests import synth  # noqa: E402


def build_reference(cfg):
    ref_shims.install()
    from models.dreamvla_model import DreamVLA  # the reference's own class
    fake = "/tmp/_dvla_fake_vit.pth"
    if not os.path.exists(fake):
        torch.save({"model": {}}, fake)
    torch.manual_seed(0)
    kw = synth.ctor_kwargs(cfg)
    model = DreamVLA(finetune_type="calvin", clip_device="cpu", vit_checkpoint_path=fake, **kw)
    model._init_model_type()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    return model


class Capture:
    """Records the tensors torch samples inside the reference forward, in call order."""

    def __init__(self):
        self.draws = []

    def __enter__(self):
        self._orig = {n: getattr(torch, n) for n in ("randn_like", "randint", "rand", "randn")}
        for n, f in self._orig.items():
            def wrap(*a, _f=f, _n=n, **k):
                if k.get("device") == "cuda":          # FMDiffusion.ddim_sample_loop hard-codes device='cuda' (respace.py:139)
                    k.pop("device")
                t = _f(*a, **k)
                self.draws.append((_n, t.clone()))
                return t
            setattr(torch, n, wrap)
        return self

    def __exit__(self, *exc):
        for n, f in self._orig.items():
            setattr(torch, n, f)


def identity_masking(x, mask_ratio):
    N, L, D = x.shape
    ids = torch.arange(L).unsqueeze(0).repeat(N, 1)
    return x, torch.zeros(N, L), ids


def run_case(name, cfg):
    model = build_reference(cfg)
    sd = synth.synth_state_dict({k: v for k, v in model.state_dict().items()}, seed=cfg["weight_seed"])
    missing = model.load_state_dict(sd, strict=True)
    inputs = synth.synth_inputs(cfg)
    orig_masking = model.vision_encoder.random_masking
    model.vision_encoder.random_masking = identity_masking
    fixture = {"cfg": cfg, "name": name}
    tensors = {}
    # ---- train mode ----
    model.train()
    torch.manual_seed(cfg["draw_seed"])
    with Capture() as cap, torch.no_grad():
        out = model(inputs["image_primary"], inputs["image_wrist"], inputs["state"], inputs["text_token"], action=None,
                    action_label=inputs["action_label"])
    names = ["arm", "gripper", "image_pred", None, None, None, "depth_pred", "traj_pred", "dino_pred", "sam_pred"]
    if cfg["use_dit_head"]:
        kinds = [d[0] for d in cap.draws]
        assert kinds == ["randn_like", "randint", "rand"], kinds
        tensors["diffusion_noise"], tensors["diffusion_timestep"] = cap.draws[0][1], cap.draws[1][1]
        tensors["diffusion_drop_ids"] = cap.draws[2][1] < 0.1          # models.py:83
        fixture["loss_action"] = float(out[0])
    else:
        tensors["arm_sub"], tensors["gripper_sub"] = synth.subsample(out[0]), synth.subsample(out[1])
    for nm, o in zip(names, out):
        if nm in ("image_pred", "depth_pred", "traj_pred", "dino_pred", "sam_pred") and o is not None:
            tensors[nm + "_sub"] = synth.subsample(o)
            fixture[nm + "_shape"] = list(o.shape)
            fixture[nm + "_norm"] = float(o.float().norm())
    # ---- permutation invariance record (random masking back on) ----
    model.vision_encoder.random_masking = orig_masking
    torch.manual_seed(123)
    with torch.no_grad():
        fp, _, _ = model.vision_encoder.forward_encoder(inputs["image_primary"].flatten(0, 1), 0.0)
        rp = model.perceiver_resampler(fp[:, 1:, :].unsqueeze(1).unsqueeze(1))
    model.vision_encoder.random_masking = identity_masking
    with torch.no_grad():
        fp0, _, _ = model.vision_encoder.forward_encoder(inputs["image_primary"].flatten(0, 1), 0.0)
        rp0 = model.perceiver_resampler(fp0[:, 1:, :].unsqueeze(1).unsqueeze(1))
    fixture["perm_invariance_rel"] = float((rp - rp0).norm() / rp0.norm())
    fixture["perm_cls_rel"] = float((fp[:, 0] - fp0[:, 0]).norm() / fp0[:, 0].norm())
    tensors["vit_primary_sub"] = synth.subsample(fp0)
    tensors["resampler_primary_sub"] = synth.subsample(rp0)
    # ---- test mode (DDIM) ----
    if cfg["use_dit_head"]:
        model.eval()
        torch.manual_seed(cfg["draw_seed"] + 1)
        with Capture() as cap, torch.no_grad():
            out_t = model(inputs["image_primary"], inputs["image_wrist"], inputs["state"], inputs["text_token"], mode="test")
        assert cap.draws[0][0] == "randn", [d[0] for d in cap.draws][:3]
        tensors["sample_noise"] = cap.draws[0][1]
        if cfg.get("use_fm", False):               # the loop ignores that noise and draws its own start state [2 bs, T, 7]
            assert [d[0] for d in cap.draws] == ["randn", "randn"] and cap.draws[1][1].shape[0] == 2 * cap.draws[0][1].shape[0]
            tensors["sample_noise"] = cap.draws[1][1]
        tensors["test_arm"], tensors["test_gripper"] = out_t[0].clone(), out_t[1].clone()
    # ---- mask (bit-exact contract) ----
    vis = (model.attention_mask == 0)
    fixture["mask_shape"] = list(vis.shape)
    fixture["mask_visible_pairs"] = int(vis.sum())
    fixture["mask_row_counts_first_timestep"] = vis[: vis.shape[0] // cfg["sequence_length"]].sum(1).tolist()
    tensors["mask_packed"] = torch.from_numpy(__import__("numpy").packbits(vis.numpy(), axis=1))
    torch.save(tensors, os.path.join(HERE, f"{name}.pt"))
    with open(os.path.join(HERE, f"{name}.json"), "w") as f:
        json.dump(fixture, f, indent=1)
    print(name, {k: v for k, v in fixture.items() if not isinstance(v, (list, dict))})
    return model, sd, inputs, tensors


if __name__ == "__main__":
    torch.set_num_threads(8)
    only = sys.argv[1:]
    for name, cfg in {**synth.CASES, **synth.CPU_ONLY_CASES, **synth.FULL_CASES, **synth.FM_CASES}.items():
        if not only or name in only:
            run_case(name, cfg)
