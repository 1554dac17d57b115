"""Small driver for ncu: a handful of GEMM / attention launches at the workload's shapes (run under ncu on the GPU box)."""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from dreamvla_b200 import _lib as L  # noqa: E402

dev = "cuda"
what = sys.argv[1] if len(sys.argv) > 1 else "gemm"
if what == "gemm":
    for (M, N, K, act) in [(8200, 4096, 1024, 1), (2580, 3072, 1024, 0), (8200, 1024, 4096, 0)]:
        A = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        W = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
        b = torch.randn(N, device=dev, dtype=torch.bfloat16)
        R = torch.randn(M, N, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            L.gemm(A, W, bias=b, act=act, residual=R if act == 0 else None)
    torch.cuda.synchronize()
else:
    B, H, Lq = (40, 16, 265) if what == "attn_dec" else (2, 16, 1290)
    qkv = torch.randn(B, Lq, 3, H, 64, device=dev, dtype=torch.bfloat16)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    for _ in range(3):
        o, lse = L.attn_fwd(q, k, v, 0.125)
    d_o = torch.randn_like(o)
    dqkv = torch.empty_like(qkv)
    for _ in range(2):
        L.attn_bwd(q, k, v, o, d_o, lse, 0.125, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2])
    torch.cuda.synchronize()
