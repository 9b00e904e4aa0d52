"""Epilogue cost of dl_gemm on the FFN shapes: the same GEMM with a plain store, with the forward
epilogue (bias + pre-activation copy + GELU) and with the backward one (multiply by GELU'(pre)).

    python tools/gemm_epi_probe.py            # CUDA-graph timing of each variant
    python tools/gemm_epi_probe.py --once     # one launch per variant (for ncu --set full)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from druglamp_b200 import _lib  # noqa: E402

M, N, K = 16384, 1024, 256


def variants():
    bf = torch.bfloat16
    A = torch.randn(M, K, device="cuda").to(bf)
    W = torch.randn(N, K, device="cuda").to(bf)          # [N, K] K-major
    Wt = torch.randn(K, N, device="cuda").to(bf)         # [K, N] MN-major
    C = torch.empty(M, N, device="cuda", dtype=bf)
    pre = torch.randn(M, N, device="cuda").to(bf)
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").to(bf)
    base = dict(M=M, N=N, K=K, lda=K, ldc=N)
    return {
        "plain": lambda: _lib.gemm(A, W, C, ldb=K, **base),
        "bias": lambda: _lib.gemm(A, W, C, ldb=K, bias=bias, **base),
        "bias+gelu": lambda: _lib.gemm(A, W, C, ldb=K, bias=bias, act=_lib.ACT_GELU, **base),
        "bias+preact+gelu": lambda: _lib.gemm(A, W, C, ldb=K, bias=bias, act=_lib.ACT_GELU, preact_out=pre, **base),
        "residual": lambda: _lib.gemm(A, W, C, ldb=K, residual=res, **base),
        "bias+drop+residual": lambda: _lib.gemm(A, W, C, ldb=K, bias=bias, residual=res, drop_p=0.1, drop_seed=7, **base),
        "gelu_grad(tb)": lambda: _lib.gemm(A, Wt, C, ldb=N, trans_b=True, mul_aux=pre,
                                           mul_mode=_lib.MUL_GELU_GRAD, **base),
        "plain(tb)": lambda: _lib.gemm(A, Wt, C, ldb=N, trans_b=True, **base),
    }


def main():
    v = variants()
    if "--once" in sys.argv:
        for f in v.values():
            f()
        torch.cuda.synchronize()
        return
    for name, f in v.items():
        f()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                f()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 10.0
        print(f"{name:22s} {us:7.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TF/s", flush=True)


if __name__ == "__main__":
    main()
