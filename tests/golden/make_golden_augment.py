"""Golden vectors of the collator's augmentation pieces, produced by the UNMODIFIED reference code on CPU:

    python tests/golden/make_golden_augment.py        (build container only: the reference is not on the GPU box)

`utils/data_utils.py` cannot be imported here (it needs omegaconf, h5py, clip, ... at module level), so the two definitions
that matter -- `class RandomShiftsAug` (:326-383) and `def depth_image_fn` (:3588-3603) -- are cut out of the file with `ast`
and executed UNCHANGED in a namespace holding the names they use (torch, nn, F, np, torchvision.transforms as T).
`RandomShiftsAug` draws its own shifts with torch.randint; the global seed is set before each call and the same draw is
repeated afterwards (same size / dtype / bounds, data_utils.py:345-349 and :372-376) to record the integer shifts.
"""
from __future__ import annotations

import ast
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision.transforms as T

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/utils/data_utils.py"


def reference_namespace():
    src = open(REF).read()
    tree = ast.parse(src)
    want = {"RandomShiftsAug": None, "depth_image_fn": None}
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in want and want[node.name] is None:
            want[node.name] = ast.get_source_segment(src, node)
    assert all(want.values()), want
    ns = {"torch": torch, "nn": nn, "F": F, "np": np, "T": T}
    for code in want.values():
        exec(compile(code, REF, "exec"), ns)
    return ns


def main():
    ns = reference_namespace()
    out = {}
    g = torch.Generator().manual_seed(20260924)
    for name, (n, t, c, hw, pad) in {"rgb": (2, 3, 3, 40, 10), "gripper": (2, 3, 3, 28, 4), "depth": (2, 3, 1, 40, 10)}.items():
        aug = ns["RandomShiftsAug"](pad)
        x = torch.randn(n, t, c, hw, hw, generator=g)
        torch.manual_seed(11)
        y_traj = aug.forward_traj(x.clone())
        torch.manual_seed(11)
        s_traj = torch.randint(1, 2 * pad + 1, size=(n * t, 1, 1, 2), dtype=x.dtype).view(n * t, 2).to(torch.int32)
        x4 = x.view(n * t, c, hw, hw)
        torch.manual_seed(12)
        y_fwd = aug(x4.clone())
        torch.manual_seed(12)
        s_fwd = torch.randint(0, 2 * pad + 1, size=(n * t, 1, 1, 2), dtype=x.dtype).view(n * t, 2).to(torch.int32)
        out[name] = {"pad": pad, "x": x, "traj": y_traj, "traj_shifts": s_traj, "fwd": y_fwd, "fwd_shifts": s_fwd}
    # depth_image_fn: a list of (H, W) images -> NEAREST resize to 224 x 224.  The real CALVIN depth maps are 200 x 200; the
    # index map of that size is pinned with a coordinate image, the values with a small random one of the same ratio (25 : 28)
    coord = np.arange(200 * 200, dtype=np.float32).reshape(200, 200)
    lin = ns["depth_image_fn"]([coord])[0, 0].to(torch.int64)                                       # [224, 224] source linear index
    ys, xs = lin[:, 0] // 200, lin[0, :] % 200
    assert torch.equal(lin, ys.view(-1, 1) * 200 + xs.view(1, -1))                                  # separable: store the two index rows
    out["resize_rows_200_to_224"], out["resize_cols_200_to_224"] = ys.to(torch.int16), xs.to(torch.int16)
    small = [torch.randn(50, 50, generator=g).numpy() for _ in range(1)]
    out["resize_small_in"] = torch.from_numpy(np.stack(small))
    small_res = ns["depth_image_fn"](small)
    assert small_res.shape == (1, 1, 224, 224)
    out["resize_small_out"] = small_res.to(torch.float16)       # values are copies of fp32 inputs; fp16 keeps the file small and
                                                               # still identifies the source pixel (compared after the same cast)
    torch.save(out, os.path.join(HERE, "augment.pt"))
    print({k: (tuple(v.shape) if torch.is_tensor(v) else {kk: (tuple(vv.shape) if torch.is_tensor(vv) else vv) for kk, vv in v.items()})
           for k, v in out.items()})


if __name__ == "__main__":
    main()
