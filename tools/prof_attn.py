"""ncu driver: a few attention launches at the path's per-GPU-batch-8 shapes (kernel variant chosen by DVLA_ATTN_FWD/BWD)."""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from dreamvla_b200 import _lib as L  # noqa: E402

dev = "cuda"
what = sys.argv[1] if len(sys.argv) > 1 else "dec"
B, H, Lq = {"dec": (160, 16, 265), "vit": (160, 12, 197), "gpt": (8, 16, 1290)}[what]
qkv = torch.randn(B, Lq, 3, H, 64, device=dev, dtype=torch.bfloat16)
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
for _ in range(3):
    o, lse = L.attn_fwd(q, k, v, 0.125)
if len(sys.argv) > 2 and sys.argv[2] == "bwd":
    d_o = torch.randn_like(o)
    dqkv = torch.empty_like(qkv)
    for _ in range(2):
        L.attn_bwd(q, k, v, o, d_o, lse, 0.125, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2])
torch.cuda.synchronize()
