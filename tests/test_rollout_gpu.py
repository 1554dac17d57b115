"""Rollout wrappers on the GPU: window / padding / row-selection semantics of `ModelWrapper.step`
(reference utils/eval_utils_calvin.py:107-145, utils/eval_utils_libero.py:94-179) against direct `mode='test'` forwards,
and rollout-level incremental inference against the full-window path."""
import numpy as np
import pytest
import torch

from tests import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def build(cfg, dev):
    from dreamvla_b200.models import DreamVLA
    torch.manual_seed(0)
    m = DreamVLA(finetune_type="calvin", clip_device="cpu", vit_checkpoint_path=None, **synth.ctor_kwargs(cfg))
    m.load_state_dict(synth.synth_state_dict(m.state_dict(), cfg["weight_seed"]))
    m = m.bfloat16().to(dev)
    m._init_model_type()
    return m.eval()


def observations(n, seed, state_dim=15):
    g = torch.Generator().manual_seed(seed)
    text = torch.zeros(77, dtype=torch.long)
    text[0], text[1:6], text[6] = 49406, torch.randint(1, 49406, (5,), generator=g), 49407
    obs = [(torch.randn(3, 224, 224, generator=g), torch.randn(3, 224, 224, generator=g), torch.randn(state_dim, generator=g) * 0.5)
           for _ in range(n)]
    return text, obs


def reference_window_action(model, history, text, noise, S, dev):
    """The reference's step, spelled out (eval_utils_calvin.py:107-145): window of the last S observations, padded by
    repeating the last one, full mode='test' forward, first predicted step of row num_step-1 (or the last row)."""
    frames = history[-S:]
    num_step = len(frames)
    frames = frames + [frames[-1]] * (S - num_step)
    ip = torch.stack([f[0] for f in frames]).unsqueeze(0).to(dev, torch.bfloat16)
    iw = torch.stack([f[1] for f in frames]).unsqueeze(0).to(dev, torch.bfloat16)
    st = torch.stack([torch.cat([f[2][:6], f[2][-1:]]) for f in frames]).unsqueeze(0).to(dev, torch.bfloat16)
    tt = text.to(dev).view(1, 1, 77).repeat(1, S, 1)
    with torch.no_grad():
        out = model(image_primary=ip, image_wrist=iw, state=st, text_token=tt, action=None, mode="test", sample_noise=noise.to(dev))
    arm, grip = out[0], out[1]
    action = torch.concat((arm[0, :, 0, :], grip[0, :, 0, :] > 0.5), dim=-1)
    action[:, -1] = (action[:, -1] - 0.5) * 2
    action = action.cpu().detach().to(dtype=torch.float16).numpy()
    return action[num_step - 1] if num_step < S else action[-1]


@pytest.mark.parametrize("graph", [False, True])
def test_modelwrapper_step_window_semantics(dev, graph):
    from dreamvla_b200.utils.eval_utils_calvin import ModelWrapper
    cfg = synth.CASES["libero_dit"]                       # 2 layers, S = 3, DiT head
    S = cfg["sequence_length"]
    model = build(cfg, dev)
    text, obs = observations(S + 3, seed=5)
    w = ModelWrapper(model, history_len=S, action_pred_steps=3, device=dev, use_cuda_graph=graph)
    g = torch.Generator().manual_seed(9)
    for i in range(len(obs)):                             # growing window (1..S), then sliding
        noise = torch.randn(S, 3, 7, generator=g)
        got = w.step(*obs[i], text, sample_noise=noise)
        want = reference_window_action(model, obs[:i + 1], text, noise, S, dev)
        assert got.dtype == np.float16 and got.shape == (7,)
        assert np.array_equal(got, want), (i, got, want)
    # a changed instruction mid-episode is ignored until reset() (:110-113)
    other = text.clone()
    other[1:6] = (other[1:6] + 17) % 49000 + 1
    noise = torch.randn(S, 3, 7, generator=g)
    a = w.step(*obs[-1], other, sample_noise=noise)
    b = reference_window_action(model, obs + [obs[-1]], text, noise, S, dev)
    assert np.array_equal(a, b)
    w.reset()
    assert len(w.img_queue) == 0 and w.text_token is None


@pytest.mark.parametrize("case,prune", [("libero_dit", False), ("libero_dit", True), ("calvin_allheads", True)])
def test_incremental_rollout_matches_full_window(dev, case, prune):
    """Per-frame token cache + selected-timestep-only backbone rows / DDIM vs the full-window forward.
    prune=False feeds the backbone the same L tokens as the full window: the selected rows see identical inputs, so the
    actions must agree to the last bit of the fp16 action the wrapper returns.  prune=True drops never-attended tokens:
    same mathematics, other KV tiling in the flash kernel -> equal up to bf16 rounding."""
    from dreamvla_b200.utils.eval_utils_calvin import ModelWrapper
    cfg = synth.CASES[case]
    S = cfg["sequence_length"]
    model = build(cfg, dev)
    text, obs = observations(S + 3, seed=11)
    full = ModelWrapper(model, history_len=S, device=dev, use_cuda_graph=False)
    inc = ModelWrapper(model, history_len=S, device=dev, use_cuda_graph=True, incremental=True, prune=prune)
    g = torch.Generator().manual_seed(4)
    worst = 0.0
    for i in range(len(obs)):
        noise = torch.randn(S, 3, 7, generator=g)
        a = full.step(*obs[i], text, sample_noise=noise).astype(np.float32)
        b = inc.step(*obs[i], text, sample_noise=noise).astype(np.float32)
        err = float(np.abs(a[:6] - b[:6]).max() / (np.abs(a[:6]).max() + 1e-6))
        worst = max(worst, err)
        if prune:
            assert err < 3e-2 and a[6] == b[6], (i, a, b)
        else:
            assert np.array_equal(a, b), (i, a, b)
    print(f"incremental[{case}, prune={prune}] worst relative action difference {worst:.3e}")


@pytest.mark.parametrize("ensembling", [False, True])
def test_libero_wrapper_gripper_width(dev, ensembling):
    """8-dim `--gripper_width` state through the LIBERO wrapper (eval_utils_libero.py:112-115) vs a direct forward; the
    incremental engine under the same wrapper agrees."""
    from dreamvla_b200.utils.eval_utils_libero import ModelWrapper, quaternion_to_euler
    cfg = synth.CPU_ONLY_CASES["libero_gripper_width"]
    S = cfg["sequence_length"]
    model = build(cfg, dev)
    g = torch.Generator().manual_seed(21)
    text, _ = observations(1, seed=3)
    w = ModelWrapper(model, history_len=S, use_ensembling=ensembling, libero_eval_max_steps=16, gripper_width=True, device=dev,
                     use_cuda_graph=False)
    wi = ModelWrapper(model, history_len=S, use_ensembling=ensembling, libero_eval_max_steps=16, gripper_width=True, device=dev,
                      use_cuda_graph=True, incremental=True, prune=False)
    w.reset()
    wi.reset()
    hist = []
    for t in range(S + 2):
        img, grip = torch.randn(3, 224, 224, generator=g), torch.randn(3, 224, 224, generator=g)
        pos, quat, qpos = torch.randn(3, generator=g).numpy() * 0.3, torch.randn(4, generator=g).numpy(), torch.rand(2, generator=g).numpy() * 0.04
        noise = torch.randn(S, 3, 7, generator=g)
        a = w.step(img, grip, pos, quat, qpos, text, t, sample_noise=noise)
        b = wi.step(img, grip, pos, quat, qpos, text, t, sample_noise=noise)
        assert a.shape == (7,) and a[-1] in (-1.0, 1.0)
        assert np.allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), atol=0, rtol=0), (t, a, b)
        st = torch.from_numpy(np.concatenate([pos, quaternion_to_euler(quat), qpos]))
        hist.append((img, grip, st))
        if not ensembling:
            frames = hist[-S:]
            n = len(frames)
            frames = frames + [frames[-1]] * (S - n)
            with torch.no_grad():
                out = model(image_primary=torch.stack([f[0] for f in frames]).unsqueeze(0).to(dev, torch.bfloat16),
                            image_wrist=torch.stack([f[1] for f in frames]).unsqueeze(0).to(dev, torch.bfloat16),
                            state=torch.stack([f[2] for f in frames]).unsqueeze(0).to(dev, torch.bfloat16),
                            text_token=text.to(dev).view(1, 1, 77).repeat(1, S, 1), mode="test", sample_noise=noise.to(dev))
            row = n - 1 if n < S else -1
            want = torch.cat((out[0][0, row, 0].float(), (out[1][0, row, 0] > 0.5).float() * 2 - 1)).to(torch.float16).cpu().numpy()
            assert np.array_equal(a, want), (t, a, want)
