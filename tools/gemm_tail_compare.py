"""K-split tail vs whole tiles on the same inputs: `dump <path>` writes the outputs of a list of GEMMs (run it once with
DVLA_GEMM_TAIL=0 and once with the default), `compare <a> <b>` reports the difference.  Both accumulate in fp32 and round once,
so the outputs may differ only through the fp32 summation order: at most one bf16 ulp on a few elements."""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])

SHAPES = [  # (M, N, K, b_mn, bias, act, residual): the small GEMMs of the B=1 parity cases and of action inference
    (21, 768, 3072, False, True, 0, True), (21, 3072, 768, False, True, 2, False), (21, 768, 768, False, True, 0, False),
    (21, 2304, 768, False, True, 0, False), (273, 1024, 4096, True, True, 0, True), (273, 4096, 1024, True, True, 2, False),
    (273, 3072, 1024, True, True, 0, False), (273, 1024, 1024, True, True, 0, True), (1290, 1024, 4096, True, True, 0, True),
    (1290, 4096, 1024, True, True, 2, False), (2580, 1024, 4096, False, False, 0, False), (1379, 768, 3072, False, True, 1, True),
    (197, 768, 3072, False, True, 0, True), (12, 768, 3072, False, True, 0, False), (640, 512, 2048, False, False, 0, False),
    (10320, 1024, 4096, True, True, 0, True), (10320, 1024, 3072, False, False, 0, False), (31520, 768, 3072, False, True, 0, True),
]


def run():
    from dreamvla_b200 import _lib as L
    g = torch.Generator().manual_seed(3)
    outs = []
    for (M, N, K, b_mn, hb, act, hr) in SHAPES:
        A = (torch.randn(M, K, generator=g) * 0.5).cuda().bfloat16()
        W = (torch.randn(N, K, generator=g) * 0.05).cuda().bfloat16()
        w = W.t().contiguous() if b_mn else W
        bias = torch.randn(N, generator=g).cuda().bfloat16() if hb else None
        res = torch.randn(M, N, generator=g).cuda().bfloat16() if hr else None
        o1 = L.gemm(A, w, b_mn=b_mn, bias=bias, act=act, residual=res)
        o2 = L.gemm(A, w, b_mn=b_mn, bias=bias, act=act, residual=res)
        ref = A.float() @ W.float().t()
        outs.append({"shape": (M, N, K, b_mn, hb, act, hr), "out": o1.cpu(), "again": o2.cpu(),
                     "plain_err": float(((L.gemm(A, w, b_mn=b_mn).float() - ref).norm() / ref.norm()))})
    return outs


if __name__ == "__main__":
    if sys.argv[1] == "dump":
        torch.save(run(), sys.argv[2])
    else:
        a, b = torch.load(sys.argv[2]), torch.load(sys.argv[3])
        bad = 0
        for x, y in zip(a, b):
            d = (x["out"].float() - y["out"].float())
            scale = float(x["out"].float().abs().max())
            ndiff = int((d != 0).sum())
            rel = float(d.norm() / x["out"].float().norm())
            again = float((y["out"].float() - y["again"].float()).abs().max())
            ok = float(d.abs().max()) <= scale * 2.0 ** -7 and rel < 5e-4 and again == 0.0
            bad += not ok
            print(f"{'PASS' if ok else 'FAIL'} {str(x['shape']):48s} differing {ndiff:8d} / {d.numel():9d}  max|d| {float(d.abs().max()):.3e} "
                  f"(max|out| {scale:.2f})  rel-L2 {rel:.2e}  repeat max|d| {again:.1e}  err vs fp32 {x['plain_err']:.3e} / {y['plain_err']:.3e}")
        print("TAIL_COMPARE", "PASS" if not bad else f"FAIL ({bad})")
