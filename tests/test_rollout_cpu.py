"""Host-side logic of the rollout wrappers (no GPU): the pruned-token closure of DreamVLA._rollout_mask against the
reference mask rules, the LIBERO state / ensembling arithmetic against the reference's expressions
(utils/eval_utils_libero.py:30-34,106-115,159-176)."""
import numpy as np
import pytest
import torch

from dreamvla_b200.models.dreamvla_model import generate_attention_mask
from dreamvla_b200.utils.eval_utils_libero import quaternion_to_euler


def closure(vis, sel, n_tok):
    keep = torch.zeros(vis.shape[0], dtype=torch.bool)
    keep[sel * n_tok:(sel + 1) * n_tok] = True
    while True:
        grown = keep | vis[keep].any(dim=0)
        if bool((grown == keep).all()):
            return keep
        keep = grown


@pytest.mark.parametrize("S,n_obs,kw", [
    (10, 54, {}),                                                       # eval.sh: obs + depth + sam queries
    (10, 90, {}),                                                       # C2: five heads
    (7, 0, {}),                                                         # LIBERO, world heads off
    (5, 18, dict(atten_goal=4, atten_goal_state=True, atten_only_obs=True, attn_robot_proprio_state=True)),
])
def test_pruned_rollout_tokens_are_exactly_the_unattended_ones(S, n_obs, kw):
    n_a, act = 36, 3
    args = dict(atten_goal=0, atten_goal_state=False, atten_only_obs=False, attn_robot_proprio_state=False, mask_l_obs_ratio=0.0)
    args.update(kw)
    m = generate_attention_mask(K=S, num_A=n_a, num_B=n_obs + act, num_obs_token=n_obs, action_pred_steps=act, **args)
    vis = m == 0
    n_tok = n_a + n_obs + act
    for sel in range(S):
        keep = closure(vis, sel, n_tok)
        # no kept row sees a dropped column => softmax over the kept columns is the full softmax
        assert not bool(vis[keep][:, ~keep].any())
        if not kw:      # plain finetune mask: A slots of timesteps <= sel and the B slots of sel, nothing else
            want = torch.zeros_like(keep)
            for t in range(sel + 1):
                want[t * n_tok:t * n_tok + n_a] = True
            want[sel * n_tok:(sel + 1) * n_tok] = True
            assert torch.equal(keep, want)
            assert int(keep.sum()) == n_a * (sel + 1) + n_obs + act


def test_quaternion_to_euler_is_scipy_xyz():
    R = pytest.importorskip("scipy.spatial.transform").Rotation
    g = np.random.default_rng(0)
    for _ in range(50):
        q = g.normal(size=4)
        assert np.allclose(quaternion_to_euler(q), R.from_quat(q).as_euler("xyz", degrees=False), atol=1e-9)


def test_libero_ensembling_matches_reference_expression():
    """The temporal-ensembling arithmetic of utils/eval_utils_libero.py:159-176, restated line by line here as the checker,
    against ModelWrapper.step driven with canned per-step predictions (no model, CPU tensors)."""
    from dreamvla_b200.utils.eval_utils_libero import ModelWrapper
    steps, T, temp = 3, 12, 0.01
    g = torch.Generator().manual_seed(1)
    preds = [(torch.randn(steps, 6, generator=g), torch.rand(steps, 1, generator=g)) for _ in range(T)]

    class Canned(ModelWrapper):
        def infer(self, *a, **k):
            self._i = getattr(self, "_i", -1) + 1
            return preds[self._i][0], preds[self._i][1], min(self._i + 1, self.history_len)
    w = Canned(torch.nn.Identity(), history_len=7, use_ensembling=True, ensembling_temp=temp, libero_eval_max_steps=T,
               action_pred_steps=steps, device="cpu", use_cuda_graph=False)
    w.reset()
    w._i = -1
    all_time = torch.zeros(T, T + steps, 7)
    z3, z4, z2, img, txt = np.zeros(3), np.array([0, 0, 0, 1.0]), np.zeros(2), torch.zeros(3, 224, 224), torch.zeros(77, dtype=torch.long)
    for t in range(T):
        got = w.step(img, img, z3, z4, z2, txt, t)
        action = torch.cat(preds[t], dim=-1).unsqueeze(0)
        all_time[t:t + 1, t:t + steps] = action
        cur = all_time[:, t]
        cur = cur[torch.all(cur != 0, axis=1)]
        ew = np.exp(-temp * np.arange(len(cur)))
        ew = torch.from_numpy(ew / ew.sum()).unsqueeze(dim=1)
        a = (cur * ew).sum(dim=0, keepdim=True)
        a = torch.concat((a[:, :6], a[:, 6:] > 0.5), dim=-1)
        a[:, -1] = (a[:, -1] - 0.5) * 2
        want = a.numpy()[-1]
        assert np.allclose(got, want, atol=1e-7), (t, got, want)
        assert got[-1] in (-1.0, 1.0) and float(w.gripper_state[0]) == got[-1]


def test_libero_state_layout():
    from dreamvla_b200.utils.eval_utils_libero import ModelWrapper
    w = ModelWrapper(torch.nn.Identity(), gripper_width=True, device="cpu", use_cuda_graph=False)
    s = w.build_state([0.1, 0.2, 0.3], [0, 0, 0, 1.0], [0.02, -0.02])
    assert s.shape == (8,) and np.allclose(s[:3], [0.1, 0.2, 0.3]) and np.allclose(s[3:6], 0) and np.allclose(s[6:], [0.02, -0.02])
    w2 = ModelWrapper(torch.nn.Identity(), gripper_width=False, device="cpu", use_cuda_graph=False)
    s2 = w2.build_state([0.1, 0.2, 0.3], [0, 0, 0, 1.0])
    assert s2.shape == (7,) and s2[-1] == -1.0                         # initial gripper command (:84)
    assert w.state_dim == 8 and w2.state_dim == 7
