"""The two largest GEMMs of the step (MHLA lin1, model/PMMA/encoder.py:128): 16384 x 2048 x 512 with the
forward epilogue (bias + GELU + stored derivative) and its dX twin (x stored derivative), each against the same
GEMM with a plain store -- graph timing, or one launch per variant for ncu.

    python tools/mhla_gemm_probe.py [--once]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from druglamp_b200 import _lib  # noqa: E402

M, N, K = 16384, 2048, 512


def variants():
    bf = torch.bfloat16
    A = torch.randn(M, K, device="cuda").to(bf)
    W = torch.randn(N, K, device="cuda").to(bf)
    C = torch.empty(M, N, device="cuda", dtype=bf)
    pre = torch.randn(M, N, device="cuda").to(bf)
    bias = torch.randn(N, device="cuda")
    G = torch.randn(M, 8, device="cuda").to(bf)          # dlogits of the 8 heads
    W2 = torch.randn(8, N, device="cuda").to(bf)         # lin2 weight [8, 2048]: MN-major B of the dX GEMM
    base = dict(M=M, N=N, K=K, lda=K, ldc=N)
    return {
        "plain": lambda: _lib.gemm(A, W, C, ldb=K, **base),
        "bias+gelu": lambda: _lib.gemm(A, W, C, ldb=K, bias=bias, act=_lib.ACT_GELU, **base),
        "bias+gelu+deriv (forward)": lambda: _lib.gemm(A, W, C, ldb=K, bias=bias, act=_lib.ACT_GELU, preact_out=pre, pre_mode=1, **base),
        "K=8 x aux (dpre of lin2)": lambda: _lib.gemm(G, W2, C, M=M, N=N, K=8, lda=8, ldb=N, ldc=N, trans_b=True, mul_aux=pre,
                                                      mul_mode=_lib.MUL_VALUE),
    }


def main():
    v = variants()
    if "--once" in sys.argv:
        for f in v.values():
            f(); f()
        torch.cuda.synchronize()
        return
    for name, f in v.items():
        f()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10):
                f()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / 50
        print(f"{name:32s} {us:7.1f} us")


if __name__ == "__main__":
    main()
