"""Pins the CPU oracle (oracle/dreamvla_oracle.py): (1) against the golden vectors generated from the unmodified
reference (tests/golden/*.pt, everywhere), (2) against the live reference modules when /root/reference is present."""
import json
import os

import pytest
import torch

from oracle import dreamvla_oracle as O
from oracle import ref_shims
from tests import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 2e-5   # fp32 vs fp32, different summation orders


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))


def template_state(cfg):
    """Shapes of every state_dict entry the oracle reads, derived from the fixture's config (no reference needed)."""
    from tests.state_template import build_template
    return build_template(cfg)


@pytest.mark.parametrize("name", list(synth.CASES) + list(synth.CPU_ONLY_CASES))
def test_oracle_matches_reference_golden(name):
    cfg = {**synth.CASES, **synth.CPU_ONLY_CASES}[name]
    fx = json.load(open(os.path.join(GOLDEN, f"{name}.json")))
    gold = torch.load(os.path.join(GOLDEN, f"{name}.pt"))
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    sd = synth.synth_state_dict(template_state(cfg), cfg["weight_seed"])
    inp = synth.synth_inputs(cfg)
    with torch.no_grad():
        out = O.dreamvla_forward(sd, cfg, inp["image_primary"], inp["image_wrist"], inp["state"], inp["text_token"],
                                 action_label=inp["action_label"], mode="train",
                                 diffusion_noise=gold.get("diffusion_noise"),
                                 diffusion_timestep=gold.get("diffusion_timestep"),
                                 diffusion_drop_ids=gold.get("diffusion_drop_ids"))
    for k in ("image_pred", "depth_pred", "traj_pred", "dino_pred", "sam_pred"):
        if k + "_sub" in gold:
            assert list(out[k].shape) == fx[k + "_shape"]
            assert rel(synth.subsample(out[k]), gold[k + "_sub"]) < TOL, k
    assert rel(synth.subsample(out["vit_primary"]), gold["vit_primary_sub"]) < TOL
    assert rel(synth.subsample(out["resampler_primary"]), gold["resampler_primary_sub"]) < TOL
    if cfg["use_dit_head"]:
        assert abs(float(out["loss_action"]) - fx["loss_action"]) < 1e-4 * abs(fx["loss_action"])
        with torch.no_grad():
            o2 = O.dreamvla_forward(sd, cfg, inp["image_primary"], inp["image_wrist"], inp["state"], inp["text_token"],
                                    mode="test", sample_noise=gold["sample_noise"])
        assert rel(o2["arm_pred_action"], gold["test_arm"]) < TOL
        assert rel(o2["gripper_pred_action"], gold["test_gripper"]) < TOL
    else:
        assert rel(synth.subsample(out["arm_pred_action"]), gold["arm_sub"]) < TOL
        assert rel(synth.subsample(out["gripper_pred_action"]), gold["gripper_sub"]) < TOL
    # the ViT token permutation of the reference (vit_mae.py:157-182) is output-invariant at fp32 noise level
    assert fx["perm_invariance_rel"] < 5e-6 and fx["perm_cls_rel"] < 5e-6


@pytest.mark.skipif(not ref_shims.reference_available(), reason="reference tree not present (GPU box)")
def test_state_template_matches_reference_keys():
    """The shape template used off-box has exactly the reference's state_dict entries (minus the dead MAE decoder)."""
    import sys
    sys.path.insert(0, GOLDEN)
    import make_golden
    cfg = synth.CASES["calvin_allheads"]
    model = make_golden.build_reference(cfg)
    ref = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    tmpl = {k: tuple(v.shape) for k, v in template_state(cfg).items()}
    dead = [k for k in ref if k.startswith("vision_encoder.decoder") or k == "vision_encoder.mask_token"]
    for k in dead:
        ref.pop(k)
    assert ref == tmpl
