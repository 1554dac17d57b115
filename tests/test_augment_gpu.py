"""Device-side collator pieces through the C ABI: dvla_shift_crop against the oracle's closed form (exact --
it is a gather) and against the reference's own RandomShiftsAug outputs (tests/golden/augment.pt; the reference's fp32 grid
arithmetic lands ~1e-5 pixels off the pixel centres, tolerance 2e-4 of the value range, written below), dvla_resize_nearest
against torchvision NEAREST (exact), the `RandomShiftsAug` / `depth_image_fn` / `augment_batch` mirrors of
dreamvla_b200/utils/data_utils.py."""
import os

import pytest
import torch

from oracle import augment_oracle as ao

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "augment.pt")


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD)


@pytest.mark.parametrize("name", ["rgb", "gripper", "depth"])
def test_shift_crop_matches_oracle_and_reference(dev, gold, name):
    from dreamvla_b200 import _lib as L
    g = gold[name]
    x, pad = g["x"], g["pad"]
    n, t = x.shape[:2]
    x4 = x.view(n * t, *x.shape[2:]).contiguous()
    for shifts, ref in ((g["traj_shifts"], g["traj"].view_as(x4)), (g["fwd_shifts"], g["fwd"])):
        want = ao.shift_crop(x4, shifts, pad)
        got = L.shift_crop(x4.to(dev), shifts.to(dev), pad)
        assert torch.equal(got.cpu(), want), "fp32 -> fp32 crop is a gather: must be exact"
        got16 = L.shift_crop(x4.to(dev), shifts.to(dev), pad, out_dtype=torch.bfloat16)
        assert torch.equal(got16.cpu(), want.to(torch.bfloat16))
        got1616 = L.shift_crop(x4.to(dev).bfloat16(), shifts.to(dev), pad)
        assert torch.equal(got1616.cpu(), want.to(torch.bfloat16)), "crop commutes with the bf16 cast"
        assert float((got.cpu() - ref).abs().max()) <= 2e-4 * float(x.abs().max()), "vs the reference's grid_sample output"


def test_shift_crop_edge_cases(dev):
    from dreamvla_b200 import _lib as L
    x = torch.randn(3, 2, 17, 17)
    for pad, shifts in ((0, [[0, 0]] * 3), (4, [[0, 0], [8, 8], [4, 4]]), (10, [[20, 0], [0, 20], [13, 7]])):
        s = torch.tensor(shifts, dtype=torch.int32)
        assert torch.equal(L.shift_crop(x.to(dev), s.to(dev), pad).cpu(), ao.shift_crop(x, s, pad))
    s = torch.tensor([[4, 4]] * 3, dtype=torch.int32)
    assert torch.equal(L.shift_crop(x.to(dev), s.to(dev), 4).cpu(), x), "shift == pad is the identity"
    empty = L.shift_crop(torch.empty(0, 3, 8, 8, device=dev), torch.empty(0, 2, dtype=torch.int32, device=dev), 2)
    assert empty.shape == (0, 3, 8, 8)


def test_resize_nearest_matches_torchvision(dev, gold):
    from dreamvla_b200 import _lib as L
    from dreamvla_b200.utils.data_utils import depth_image_fn
    coord = torch.arange(200 * 200, dtype=torch.float32).view(1, 200, 200)
    got = L.resize_nearest(coord.to(dev), 224, 224).cpu()[0].long()
    ys, xs = gold["resize_rows_200_to_224"].long(), gold["resize_cols_200_to_224"].long()
    assert torch.equal(got, ys.view(-1, 1) * 200 + xs.view(1, -1))
    small = depth_image_fn(gold["resize_small_in"].to(dev)).cpu()
    assert small.shape == (1, 1, 224, 224) and torch.equal(small.to(torch.float16), gold["resize_small_out"])
    for (hin, win, hout, wout) in ((200, 200, 224, 224), (84, 84, 224, 224), (300, 173, 224, 224), (224, 224, 224, 224), (7, 5, 3, 11)):
        d = torch.randn(2, hin, win)
        assert torch.equal(L.resize_nearest(d.to(dev), hout, wout).cpu(), ao.resize_nearest(d, hout, wout)), (hin, win, hout, wout)
        assert torch.equal(ao.resize_nearest(d, hout, wout), torch.nn.functional.interpolate(d[None], size=(hout, wout), mode="nearest")[0])


def test_random_shifts_aug_module_and_batch(dev):
    """`RandomShiftsAug` / `augment_batch`: shifts in the reference's ranges, forward_traj shifts every frame on its own, image
    and depth windows draw independently (collator :1340-1341), a C2-sized window in one pass."""
    from dreamvla_b200.utils.data_utils import RandomShiftsAug, augment_batch
    aug = RandomShiftsAug(10)
    g = torch.Generator(device=dev).manual_seed(3)
    s = aug.draw(4000, dev, 1, g)
    assert int(s.min()) == 1 and int(s.max()) == 20
    s0 = aug.draw(4000, dev, 0, g)
    assert int(s0.min()) == 0 and int(s0.max()) == 20
    x = torch.randn(2, 13, 3, 224, 224, device=dev).bfloat16()
    shifts = aug.draw(26, dev, 1, g)
    y = aug.forward_traj(x, shifts=shifts)
    want = ao.shift_crop(x.float().cpu().view(26, 3, 224, 224), shifts.cpu(), 10).view_as(x).to(torch.bfloat16)
    assert torch.equal(y.cpu(), want)
    batch = {"images_primary": x, "images_wrist": torch.randn(2, 13, 3, 224, 224, device=dev).bfloat16(),
             "depth_primary": torch.rand(2, 13, 1, 224, 224, device=dev).bfloat16(), "states": torch.zeros(2, 13, 15, device=dev)}
    out = augment_batch(batch, rgb_pad=10, gripper_pad=4, traj_cons=True, generator=g)
    assert out["images_primary"].shape == x.shape and out["depth_primary"].shape == batch["depth_primary"].shape
    assert out["states"] is batch["states"] and not torch.equal(out["images_primary"], x)
    out2 = augment_batch(batch, rgb_pad=10, gripper_pad=-1, traj_cons=False, generator=g)
    assert out2["images_wrist"] is batch["images_wrist"] and out2["depth_primary"] is batch["depth_primary"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    big = torch.randn(8, 13, 3, 224, 224, device=dev)
    sh = aug.draw(104, dev, 1, g)
    for _ in range(3):
        aug.forward_traj(big, shifts=sh, out_dtype=torch.bfloat16)
    e0.record()
    for _ in range(10):
        aug.forward_traj(big, shifts=sh, out_dtype=torch.bfloat16)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"\nshift_crop [8,13,3,224,224] fp32 -> bf16: {us:.1f} us, {big.numel() * 6 / us / 1e3:.0f} GB/s")
