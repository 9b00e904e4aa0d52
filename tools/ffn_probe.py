"""Fused FFN kernels at the model's shape (16384 x 256 -> 1024 -> 256): event timing, and the launches an
`ncu -k regex:ffn_chain` capture picks up.

    python tools/ffn_probe.py                 # timings of dl_ffn_fwd / dl_ffn_bwd and the two-GEMM paths
    ncu --set full --clock-control none --import-source on -k regex:ffn_chain -s 4 -c 2 -o gpurun_out/ffn python tools/ffn_probe.py --once
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from druglamp_b200 import kernels as K  # noqa: E402


def main():
    once = "--once" in sys.argv
    M, D, Dh = 16384, 256, 1024
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g, device="cuda")
    x, res, g2 = r(M, D).bfloat16(), r(M, D).bfloat16(), r(M, D).bfloat16()
    w1, w2 = (r(Dh, D) / 16).bfloat16(), (r(D, Dh) / 32).bfloat16()
    b1, b2 = r(Dh) * 0.1, r(D) * 0.1
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    drop = (0.1, 11, 12)

    def fwd():
        return K.ffn_fwd(x, w1, b1, w2, b2, res, drop, keep=True)

    y, hd, dact = fwd()

    def bwd():
        return K.ffn_bwd(g2, w1, w2, dact)

    def fwd2():
        h = torch.empty_like(hd); d = torch.empty_like(hd)
        K.mm(x, w1, h, bias=b1, act=K.ACT_GELU, pre=d, drop=(0.1, 11), pre_mode=1)
        return K.mm(h, w2, bias=b2, res=res, drop=(0.1, 12))

    def bwd2():
        dp = K.mm(g2, w2, tb=True, mul_aux=dact, mul_mode=K.MUL_VALUE)
        return K.mm(dp, w1, tb=True)

    def fwd_inf():
        return K.ffn_fwd(x, w1, b1, w2, b2, res, (0.0, 0, 0), keep=False)

    for f in (fwd, bwd):
        f()
    torch.cuda.synchronize()
    if once:
        for _ in range(3):
            fwd(); bwd()
        torch.cuda.synchronize()
        return
    for name, f, flops in (("ffn_fwd fused", fwd, 4.0 * M * D * Dh), ("ffn_fwd two GEMMs", fwd2, 4.0 * M * D * Dh),
                           ("ffn_bwd fused", bwd, 4.0 * M * D * Dh), ("ffn_bwd two GEMMs", bwd2, 4.0 * M * D * Dh),
                           ("ffn_fwd fused, forward-only", fwd_inf, 4.0 * M * D * Dh)):
        ts = []
        for _ in range(12):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts = sorted(ts[2:])
        us = ts[len(ts) // 2]
        print(f"{name:32s} {us:7.1f} us   {flops / us / 1e6:7.1f} TFLOP/s")


if __name__ == "__main__":
    main()
