"""Whole-model parity on the GPU: dreamvla_b200 (CUDA kernels, bf16) vs the CPU oracle (fp32) on the golden cases.

Tolerance (BASELINE.md §4 / SURVEY §7): relative L2 error of bf16 outputs against the fp32 oracle; bf16 has 8 bits of
mantissa, so element-wise 1e-3 is not attainable through 24+ layers -- the bar used here is rel-L2 <= 2e-2 for head
outputs / losses, and bit-exact agreement for the mask and the token layout."""
import json
import os

import numpy as np
import pytest
import torch

from tests import synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


def build(cfg, dev, depth_bias=0.0):
    from dreamvla_b200.models import DreamVLA
    torch.manual_seed(0)
    m = DreamVLA(finetune_type="calvin", clip_device="cpu", vit_checkpoint_path=None, **synth.ctor_kwargs(cfg))
    sd = synth.synth_state_dict(m.state_dict(), cfg["weight_seed"])
    if depth_bias and "depth_decoder_pred.bias" in sd:
        sd["depth_decoder_pred.bias"] = sd["depth_decoder_pred.bias"] + depth_bias
    m.load_state_dict(sd)
    m = m.bfloat16().to(dev)
    m._init_model_type()
    gp = m.transformer_backbone
    gp.embd_pdrop = 0.0
    for blk in gp.h:
        blk.attn.attn_pdrop = blk.attn.resid_pdrop = blk.mlp.resid_pdrop = 0.0
    return m, sd


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


@pytest.mark.parametrize("name", list(synth.CASES) + list(synth.CPU_ONLY_CASES) + list(synth.FM_CASES))
def test_forward_matches_oracle_and_golden(name, dev):
    from oracle import dreamvla_oracle as O
    cfg = {**synth.CASES, **synth.CPU_ONLY_CASES, **synth.FM_CASES}[name]
    fx = json.load(open(os.path.join(GOLDEN, f"{name}.json")))
    gold = torch.load(os.path.join(GOLDEN, f"{name}.pt"))
    model, sd = build(cfg, dev)
    # bit-exact mask / token layout contract
    vis = (model.attention_mask.float().cpu() == 0)
    assert torch.equal(torch.from_numpy(np.packbits(vis.numpy(), axis=1)), gold["mask_packed"])
    assert int(vis.sum()) == fx["mask_visible_pairs"]
    inp = {k: v.to(dev) for k, v in synth.synth_inputs(cfg).items()}
    model.train()
    kw = {}
    if cfg["use_dit_head"]:
        kw = dict(diffusion_noise=gold["diffusion_noise"].to(dev), diffusion_timestep=gold["diffusion_timestep"].to(dev),
                  diffusion_drop_ids=gold["diffusion_drop_ids"].to(dev).long())
    with torch.no_grad():
        out = model(inp["image_primary"], inp["image_wrist"], inp["state"], inp["text_token"],
                    action_label=inp["action_label"], **kw)
    names = ["arm", "gripper", "image_pred", None, None, None, "depth_pred", "traj_pred", "dino_pred", "sam_pred"]
    for nm, o in zip(names, out):
        if nm and nm.endswith("_pred") and o is not None:
            assert list(o.shape) == fx[nm + "_shape"]
            assert rel(synth.subsample(o), gold[nm + "_sub"]) < 2e-2, nm
    if cfg["use_dit_head"]:
        assert abs(float(out[0]) - fx["loss_action"]) < 2e-2 * abs(fx["loss_action"])
        model.eval()
        with torch.no_grad():
            ot = model(inp["image_primary"], inp["image_wrist"], inp["state"], inp["text_token"], mode="test",
                       sample_noise=gold["sample_noise"].to(dev))
        assert rel(ot[0], gold["test_arm"]) < 5e-2
        assert rel(ot[1], gold["test_gripper"]) < 5e-2
    else:
        assert rel(synth.subsample(out[0]), gold["arm_sub"]) < 2e-2
        assert rel(synth.subsample(out[1]), gold["gripper_sub"]) < 2e-2


def test_flow_matching_head_backward(dev):
    """`--use_fm` (ActionModelFM): gradients of the flow-matching loss vs autograd through the fp32 oracle, same draws."""
    from oracle import dreamvla_oracle as O
    name = "libero_fm"
    cfg = synth.FM_CASES[name]
    gold = torch.load(os.path.join(GOLDEN, f"{name}.pt"))
    model, sd = build(cfg, dev)
    from dreamvla_b200.models.action_model import ActionModelFM, FMDiffusion
    assert isinstance(model.action_model, ActionModelFM) and model.action_model.diffusion.num_timesteps == 10
    assert isinstance(model.action_model.create_ddim(10), FMDiffusion)
    model.train()
    inp = synth.synth_inputs(cfg)
    dinp = {k: v.to(dev) for k, v in inp.items()}
    out = model(dinp["image_primary"], dinp["image_wrist"], dinp["state"], dinp["text_token"], action_label=dinp["action_label"],
                diffusion_noise=gold["diffusion_noise"].to(dev), diffusion_timestep=gold["diffusion_timestep"].to(dev),
                diffusion_drop_ids=gold["diffusion_drop_ids"].to(dev).long())
    out[0].float().backward()
    frozen = ("vision_encoder.", "clip_model.", "attention_mask", "position_embedding")
    osd = {k: v.clone() for k, v in sd.items()}
    for k, v in osd.items():
        if v.is_floating_point() and not any(f in k for f in frozen) or k == "transformer_backbone_position_embedding":
            v.requires_grad_(True)
    fwd = O.dreamvla_forward(osd, cfg, inp["image_primary"], inp["image_wrist"], inp["state"], inp["text_token"],
                             action_label=inp["action_label"], diffusion_noise=gold["diffusion_noise"],
                             diffusion_timestep=gold["diffusion_timestep"], diffusion_drop_ids=gold["diffusion_drop_ids"])
    fwd["loss_action"].backward()
    assert abs(float(out[0]) - float(fwd["loss_action"])) < 2e-2 * abs(float(fwd["loss_action"]))
    params = dict(model.named_parameters())
    worst = [(rel(params[k].grad, osd[k].grad), k) for k in (
        "action_model.net.blocks.0.attn.qkv.weight", "action_model.net.blocks.11.mlp.fc2.weight",
        "action_model.net.final_layer.linear.weight", "action_model.net.t_embedder.mlp.0.weight",
        "action_model.net.x_embedder.linear.weight", "action_pred_token", "transformer_backbone.h.1.mlp.c_proj.weight")]
    bad = [(e, k) for e, k in worst if not e < 6e-2]
    assert not bad, f"gradient mismatch: {bad}"


def test_backward_matches_oracle(dev):
    """Gradients of the total train loss w.r.t. a spread of parameters vs autograd through the fp32 oracle."""
    from dreamvla_b200.utils.train_utils import StepConfig, build_labels, compute_losses
    from oracle import dreamvla_oracle as O
    name = "calvin_allheads"
    cfg = synth.CASES[name]
    gold = torch.load(os.path.join(GOLDEN, f"{name}.pt"))
    # SiLog's gradient is ~1/pred: with zero-mean synthetic weights half the ReLU'd depth predictions sit at 0 and the
    # bf16-vs-fp32 comparison is dominated by ReLU sign flips; shift the depth head positive so the loss is well conditioned
    model, sd = build(cfg, dev, depth_bias=6.0)
    model.train()
    S = cfg["sequence_length"]
    inp = synth.synth_inputs(cfg)
    lab = synth.synth_labels(cfg)
    scfg = StepConfig(sequence_length=S, future_steps=3, use_dit_head=True, loss_image=True, loss_depth=True,
                      loss_dino_feat=True, loss_sam_feat=True, loss_trajectory=True, flow_as_mask=True)
    dinp = {k: v.to(dev) for k, v in inp.items()}
    dlab = {k: v.to(dev, torch.bfloat16) for k, v in lab.items()}
    out = model(dinp["image_primary"], dinp["image_wrist"], dinp["state"], dinp["text_token"],
                action_label=dinp["action_label"], diffusion_noise=gold["diffusion_noise"].to(dev),
                diffusion_timestep=gold["diffusion_timestep"].to(dev),
                diffusion_drop_ids=gold["diffusion_drop_ids"].to(dev).long())
    labels = build_labels(scfg, dlab, dict(image=True, depth=True, dino=True, sam=True, traj=True))
    terms = compute_losses(scfg, out, labels, bs=cfg["batch"])
    total = sum(terms.values())
    total.backward()
    # oracle
    frozen = ("vision_encoder.", "clip_model.", "attention_mask", "position_embedding")
    osd = {k: v.clone() for k, v in sd.items()}
    for k, v in osd.items():
        if v.is_floating_point() and not any(f in k for f in frozen) or k == "transformer_backbone_position_embedding":
            v.requires_grad_(True)
    # labels are consumed in bf16 by the CUDA path; give the oracle the same rounded labels
    lab_r = {k: v.to(torch.bfloat16).float() for k, v in lab.items()}
    fwd = O.dreamvla_forward(osd, cfg, inp["image_primary"], inp["image_wrist"], inp["state"], inp["text_token"],
                             action_label=inp["action_label"], diffusion_noise=gold["diffusion_noise"],
                             diffusion_timestep=gold["diffusion_timestep"], diffusion_drop_ids=gold["diffusion_drop_ids"])
    ol = O.train_losses(dict(cfg, future_steps=3, flow_as_mask=True), fwd, lab_r)
    ol["loss"].backward()
    assert abs(float(total) - float(ol["loss"])) < 2e-2 * abs(float(ol["loss"]))
    for k_mine, k_or in (("image", "loss_image"), ("depth", "loss_depth"), ("dino", "loss_dino"), ("sam", "loss_sam"), ("traj", "loss_traj")):
        w = {"image": 0.1, "depth": 0.001, "dino": 0.01, "sam": 0.01, "traj": 0.1}[k_mine]
        assert abs(float(terms[k_mine]) - w * float(ol[k_or])) < 3e-2 * abs(w * float(ol[k_or])) + 1e-6, k_mine
    probe = ["transformer_backbone.h.0.attn.c_attn.weight", "transformer_backbone.h.1.mlp.c_proj.weight",
             "transformer_backbone.h.0.ln_1.weight", "transformer_backbone.h.1.attn.c_proj.bias",
             "perceiver_resampler.layers.0.0.to_kv.weight", "perceiver_resampler.latents", "image_primary_projector.weight",
             "image_decoder.0.attn.qkv.weight", "depth_decoder.1.mlp.fc2.weight", "dino_decoder_pred.weight",
             "sam_feat_decoder.0.norm1.weight", "traj_decoder_obs_pred_projector.bias", "obs_tokens", "action_pred_token",
             "action_model.net.blocks.0.attn.qkv.weight", "action_model.net.final_layer.linear.weight",
             "action_model.net.z_embedder.linear.weight", "text_projector.weight", "state_projector.weight",
             "embedding_layer_norm.weight", "transformer_backbone_position_embedding", "mask_token"]
    params = dict(model.named_parameters())
    worst = []
    for k in probe:
        g_mine, g_or = params[k].grad, osd[k].grad
        assert g_mine is not None and g_or is not None, k
        worst.append((rel(g_mine, g_or), k))
    bad = [(e, k) for e, k in worst if not e < 6e-2]
    assert not bad, f"gradient mismatch: {bad}\nall: {sorted(worst, reverse=True)[:8]}"
