"""Reads the DVLA_ATTN_TRACE=1 timestamps of attn_fwd_ws_kernel (cycles since CTA start, written over the LSE rows of the
256-row CTA of each (b, h)) and prints the event timeline of a few CTAs.  Run with DVLA_ATTN_FWD=ws DVLA_ATTN_TRACE=1."""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from dreamvla_b200 import _lib as L  # noqa: E402

dev = "cuda"
B, H, Lq = (160, 16, 265) if len(sys.argv) < 2 else tuple(int(x) for x in sys.argv[1:4])
qkv = torch.randn(B, Lq, 3, H, 64, device=dev, dtype=torch.bfloat16)
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
for _ in range(2):
    o, lse = L.attn_fwd(q, k, v, 0.125)
torch.cuda.synchronize()
lse = lse.view(B, H, Lq).cpu()
nt = (Lq + 63) // 64
for (b, h) in [(0, 0), (B // 2, H // 2), (B - 1, H - 1)]:
    t = lse[b, h]
    print(f"--- CTA b={b} h={h}: MMA thread: q_full {t[0]:.0f}")
    for i in range(nt):
        print(f"  tile {i}: kv_full {t[1+6*i]:.0f}  S0 issued {t[2+6*i]:.0f}  S1 issued {t[3+6*i]:.0f}  p_ready0 {t[4+6*i]:.0f}  p_ready1 {t[5+6*i]:.0f}")
    names = ["top", "s_full", "fence_after", "tmem_ld", "arrive_s_free", "computed", "pv_done", "fence_after2", "sts", "proxy_fence"]
    for w in range(2):
        base = 64 + 64 * w
        print(f"  softmax WG{w}: deltas between consecutive events")
        prev = None
        for i in range(nt):
            e = [float(x) for x in t[base + 10 * i: base + 10 * i + 10]]
            line = []
            for n, x in zip(names, e):
                if x == 0 and n in ("pv_done", "fence_after2"):
                    continue
                line.append(f"{n}+{x - prev:.0f}" if prev is not None else f"{n}@{x:.0f}")
                prev = x
            print(f"    tile {i} (@{e[0]:.0f}): " + "  ".join(line))
        print(f"    final pv_done {t[base+60]:.0f}  stored {t[base+61]:.0f}")
