"""Host logic of dreamvla_b200/utils/data_utils.py without a GPU: the C-ABI call is replaced by the oracle's closed form (what the
kernel computes, tests/test_augment_gpu.py) so that the bookkeeping around it -- which entries of the batch are augmented, with
which pad, how shifts are drawn and shared -- is checked against the reference collator's block (utils/data_utils.py:1337-1354)."""
import torch

from dreamvla_b200 import _lib as L
from dreamvla_b200.utils import data_utils as du
from oracle import augment_oracle as ao


def fake_shift_crop(calls):
    def f(x, shifts, pad, out_dtype=None):
        calls.append((tuple(x.shape), shifts.clone(), pad))
        return ao.shift_crop(x.float(), shifts, pad).to(out_dtype or x.dtype)
    return f


def batch():
    g = torch.Generator().manual_seed(1)
    return {"images_primary": torch.randn(2, 5, 3, 16, 16, generator=g), "images_wrist": torch.randn(2, 5, 3, 12, 12, generator=g),
            "depth_primary": torch.rand(2, 5, 1, 16, 16, generator=g), "depth_wrist": torch.rand(2, 5, 1, 12, 12, generator=g),
            "states": torch.zeros(2, 5, 15), "text": torch.zeros(2, 77, dtype=torch.long)}


def test_traj_cons_augments_rgb_and_depth_with_independent_draws(monkeypatch):
    calls = []
    monkeypatch.setattr(L, "shift_crop", fake_shift_crop(calls))
    b = batch()
    g = torch.Generator().manual_seed(7)
    out = du.augment_batch(b, rgb_pad=3, gripper_pad=2, traj_cons=True, generator=g)
    # four forward_traj calls in the collator's order (:1340-1341, :1349-1350): static rgb, static depth, gripper rgb, gripper depth
    assert [c[0] for c in calls] == [(10, 3, 16, 16), (10, 1, 16, 16), (10, 3, 12, 12), (10, 1, 12, 12)]
    assert [c[2] for c in calls] == [3, 3, 2, 2]
    for shape, s, pad in calls:
        assert s.shape == (10, 2) and s.dtype == torch.int32 and int(s.min()) >= 1 and int(s.max()) <= 2 * pad      # low = 1 (:372)
    assert not torch.equal(calls[0][1], calls[1][1]), "image and depth windows draw their own shifts (reference behaviour)"
    assert out["states"] is b["states"] and out["text"] is b["text"]
    for k in ("images_primary", "images_wrist", "depth_primary", "depth_wrist"):
        assert out[k].shape == b[k].shape and out[k].dtype == b[k].dtype
    want = ao.shift_crop(b["images_primary"].view(10, 3, 16, 16), calls[0][1], 3).view(2, 5, 3, 16, 16)
    assert torch.equal(out["images_primary"], want)


def test_per_image_mode_leaves_depth_alone(monkeypatch):
    calls = []
    monkeypatch.setattr(L, "shift_crop", fake_shift_crop(calls))
    b = batch()
    out = du.augment_batch(b, rgb_pad=3, gripper_pad=-1, traj_cons=False, generator=torch.Generator().manual_seed(2))
    assert [c[0] for c in calls] == [(10, 3, 16, 16)] and int(calls[0][1].min()) >= 0                                # low = 0 (:345)
    assert out["depth_primary"] is b["depth_primary"] and out["images_wrist"] is b["images_wrist"]
    assert out["images_primary"].shape == b["images_primary"].shape


def test_no_pads_is_identity(monkeypatch):
    calls = []
    monkeypatch.setattr(L, "shift_crop", fake_shift_crop(calls))
    b = batch()
    out = du.augment_batch(b, rgb_pad=-1, gripper_pad=-1, traj_cons=True)
    assert not calls and all(out[k] is b[k] for k in b)


def test_explicit_shifts_and_generator_reproducibility(monkeypatch):
    calls = []
    monkeypatch.setattr(L, "shift_crop", fake_shift_crop(calls))
    aug = du.RandomShiftsAug(4)
    x = torch.randn(3, 2, 3, 10, 10)
    s = torch.tensor([[4, 4]] * 6, dtype=torch.int32)
    assert torch.equal(aug.forward_traj(x, shifts=s), x), "shift == pad is the identity"
    a = aug.forward_traj(x, generator=torch.Generator().manual_seed(5))
    b = aug.forward_traj(x, generator=torch.Generator().manual_seed(5))
    assert torch.equal(a, b)
    d = aug.draw(2000, torch.device("cpu"), 0, torch.Generator().manual_seed(1))
    assert int(d.min()) == 0 and int(d.max()) == 8 and d.shape == (2000, 2)
