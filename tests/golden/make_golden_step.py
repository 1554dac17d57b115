"""Golden vectors of the TRAIN STEP, produced by running the UNMODIFIED reference step loop
(`/root/reference/utils/train_utils.py:59-726` `train_one_epoch_calvin`) on CPU in fp32:

    python tests/golden/make_golden_step.py          (build container only: the reference is not on the GPU box)

The reference model (through oracle/ref_shims.py), torch.optim.AdamW and the reference's own loop run three micro-batches
with `gradient_accumulation_steps=2` and `num_batches=3`, so both optimiser-step rules of train_utils.py:602-604 fire (the
accumulation boundary after micro-step 2 and the last-batch-of-the-epoch rule after micro-step 3), and the per-micro-step
`clip_grad_norm_(0.1)` on the ACCUMULATED gradient (:599-600) is exercised twice on a non-empty accumulator.

Captured, without touching the reference source:
  * the scalar handed to `loss.backward()` each micro-step (torch.Tensor.backward wrapper);
  * the total norm `clip_grad_norm_` returns each micro-step (= norm of the accumulated, not-yet-clipped gradient);
  * the accumulated, clipped gradient of a spread of named parameters at each `optimizer.step()`;
  * the loss dictionaries the loop logs through its `wandb` argument (a recording stand-in);
  * the tensors torch samples inside the forward (diffusion noise / timesteps / label-drop draws), per micro-step;
so that the oracle (tests/test_oracle_cpu.py) and the CUDA train step (tests/test_train_step_gpu.py) can be driven with
identical inputs and draws.  The learning rate is tiny (1e-7): AdamW's first update is lr * sign(g) on every one of the 550 M
parameters, and with these synthetic weights anything larger leaves the linear regime (lr = 2e-5 moved the fp32 loss from 1.6
to 7.2) while a bf16 replica rounds such an update away -- micro-step 3 could then not be compared.  Two host-only accommodations: `Tensor.cuda()` is a no-op for the duration (the loop calls it
on the track labels, :459-460) and dropout probabilities are 0 (as in make_golden.py).
"""
from __future__ import annotations

import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
from tests import synth  # noqa: E402
from tests.golden.make_golden import Capture, build_reference, identity_masking  # noqa: E402

STEP_CASES = {
    "step_calvin_accum2": dict(model="calvin_allheads", accum=2, num_batches=3, batch=1, lr=1e-7, weight_decay=1e-4,
                               flow_as_mask=True, data_seed=4321),
}

PROBE = ["transformer_backbone.h.0.attn.c_attn.weight", "transformer_backbone.h.1.mlp.c_proj.weight",
         "transformer_backbone.h.0.ln_1.weight", "transformer_backbone.h.1.attn.c_proj.bias",
         "perceiver_resampler.layers.0.0.to_kv.weight", "perceiver_resampler.latents", "image_primary_projector.weight",
         "image_decoder.0.attn.qkv.weight", "depth_decoder.1.mlp.fc2.weight", "dino_decoder_pred.weight",
         "sam_feat_decoder.0.norm1.weight", "traj_decoder_obs_pred_projector.bias", "obs_tokens", "action_pred_token",
         "action_model.net.blocks.0.attn.qkv.weight", "action_model.net.final_layer.linear.weight",
         "action_model.net.z_embedder.linear.weight", "text_projector.weight", "state_projector.weight",
         "embedding_layer_norm.weight", "transformer_backbone_position_embedding", "mask_token"]


def step_args(case, mcfg):
    """The flags the reference loop reads (arguments_utils.py names), as scripts/CALVIN_ABC_D/DreamVLA/finetune.sh sets them."""
    S = mcfg["sequence_length"]
    return types.SimpleNamespace(
        num_epochs=1, precision="fp32", rank=0, world_size=1, report_to_wandb=True, batch_size=case["batch"],
        gradient_accumulation_steps=case["accum"], sequence_length=S, window_size=S + 3, future_steps=3, atten_goal=0,
        pred_num=1, action_pred_steps=mcfg["action_pred_steps"], patch_size=16, gripper_width=False,
        use_dit_head=mcfg["use_dit_head"], loss_action=True, obs_pred=mcfg["obs_pred"], loss_image=mcfg["obs_pred"],
        depth_pred=mcfg["depth_pred"], loss_depth=mcfg["depth_pred"], dino_feat_pred=mcfg["dino_feat_pred"],
        loss_dino_feat=mcfg["dino_feat_pred"], sam_feat_pred=mcfg["sam_feat_pred"], loss_sam_feat=mcfg["sam_feat_pred"],
        trajectory_pred=mcfg["trajectory_pred"], loss_trajectory=mcfg["trajectory_pred"], flow_as_mask=case["flow_as_mask"],
        loss_arm_action_ratio=1.0, loss_gripper_action_ratio=0.01, use_dpt_head=False, no_unshuffle=False,
        no_pred_gripper_traj=False, track_label_patch_size=mcfg["track_label_patch_size"], learning_rate=case["lr"],
        weight_decay=case["weight_decay"])


def host_batches(case, mcfg):
    """The collator's 13-tuples (data_utils.py:1395-1397) from the package's own seeded synthetic generator (fp32, CPU)."""
    from dreamvla_b200.utils.train_utils import StepConfig, synthetic_batch
    scfg = StepConfig(sequence_length=mcfg["sequence_length"], future_steps=3)
    heads = dict(depth=True, dino=True, sam=True, traj=True, flow_mask=True)
    out = []
    for i in range(case["num_batches"]):
        b = synthetic_batch(scfg, case["batch"], "cpu", seed=case["data_seed"] + i, heads=heads, dtype=torch.float32)
        out.append(b)
    return out


def as_tuple(b):
    tracks = {"tracks": b["tracks"].clone(), "tracks_gripper": b["tracks_gripper"].clone()}
    return (b["images_primary"].clone(), b["text"].clone(), b["actions"].clone(), b["images_wrist"].clone(), b["states"].clone(),
            None, b["depth_primary"].clone(), b["depth_wrist"].clone(), b["dino_primary"].clone(), b["dino_wrist"].clone(),
            b["sam_primary"].clone(), b["sam_wrist"].clone(), tracks)


class Loader:
    def __init__(self, batches):
        self.batches, self.num_batches = batches, len(batches)

    def __iter__(self):
        return iter(as_tuple(b) for b in self.batches)


class FakeWandb:
    def __init__(self):
        self.logs = []

    def log(self, d, **kw):
        self.logs.append({k: v for k, v in d.items() if isinstance(v, (int, float))})

    @staticmethod
    def Image(x, caption=None):
        return None


def run_case(name, case):
    mcfg = dict(synth.CASES[case["model"]])
    model = build_reference(mcfg)
    sd = synth.synth_state_dict(dict(model.state_dict()), seed=mcfg["weight_seed"])
    # SiLog needs pred > 0 on a useful fraction of pixels (see tests/test_model_gpu.py): shift the depth head positive
    sd["depth_decoder_pred.bias"] = sd["depth_decoder_pred.bias"] + 6.0
    model.load_state_dict(sd, strict=True)
    model.vision_encoder.random_masking = identity_masking
    model.clip_model.requires_grad_(False)
    from utils.train_utils import train_one_epoch_calvin      # the reference's own loop
    args = step_args(case, mcfg)
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=case["lr"], weight_decay=case["weight_decay"])
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0)
    named = dict(model.named_parameters())
    rec = dict(losses=[], norms=[], step_grads=[], draws=[])
    orig_backward, orig_clip, orig_step, orig_cuda = torch.Tensor.backward, torch.nn.utils.clip_grad_norm_, opt.step, torch.Tensor.cuda

    def backward(self, *a, **k):
        rec["losses"].append(float(self.detach()))
        rec["draws"].append(list(cap.draws))
        cap.draws.clear()
        return orig_backward(self, *a, **k)

    def clip(params, max_norm, *a, **k):
        assert max_norm == 0.1
        n = orig_clip(params, max_norm, *a, **k)
        rec["norms"].append(float(n))
        return n

    def step(*a, **k):
        rec["step_grads"].append({kname: synth.subsample(named[kname].grad, 2048) for kname in PROBE})
        return orig_step(*a, **k)

    wb = FakeWandb()
    loader = Loader(host_batches(case, mcfg))
    torch.manual_seed(mcfg["draw_seed"])
    torch.Tensor.backward, torch.nn.utils.clip_grad_norm_, opt.step = backward, clip, step
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        with Capture() as cap:
            train_one_epoch_calvin(args=args, model=model, epoch=0, calvin_loader=loader,
                                   optimizer=opt, lr_scheduler=sched, device_id="cpu", wandb=wb)
    finally:
        torch.Tensor.backward, torch.nn.utils.clip_grad_norm_, torch.Tensor.cuda = orig_backward, orig_clip, orig_cuda
    assert len(rec["losses"]) == case["num_batches"] and len(rec["step_grads"]) == 2, (len(rec["losses"]), len(rec["step_grads"]))
    tensors = {}
    for i, draws in enumerate(rec["draws"]):
        kinds = [d[0] for d in draws]
        assert kinds == ["randn_like", "randint", "rand"], kinds
        tensors[f"noise_{i}"], tensors[f"timestep_{i}"], tensors[f"drop_{i}"] = draws[0][1], draws[1][1], draws[2][1] < 0.1
    for j, g in enumerate(rec["step_grads"]):
        for k, v in g.items():
            tensors[f"grad{j}:{k}"] = v
    after = dict(model.named_parameters())
    for k in PROBE:
        tensors[f"param_after:{k}"] = synth.subsample(after[k].detach(), 2048)
    loss_logs = [l for l in wb.logs if "loss_calvin" in l]
    fixture = dict(name=name, case=case, micro_losses=rec["losses"], accumulated_grad_norms=rec["norms"], wandb_loss_logs=loss_logs,
                   optimizer_steps_after_micro=[2, 3], probe=PROBE, depth_bias_shift=6.0)
    torch.save(tensors, os.path.join(HERE, f"{name}.pt"))
    with open(os.path.join(HERE, f"{name}.json"), "w") as f:
        json.dump(fixture, f, indent=1)
    print(name, "losses", rec["losses"], "norms", rec["norms"])
    print(loss_logs)


if __name__ == "__main__":
    torch.set_num_threads(8)
    ref_shims.install()
    for name, case in STEP_CASES.items():
        run_case(name, case)
