"""Micro-benchmark of dl_gemm on the shapes the DrugLAMP step issues: each shape is captured in a
CUDA graph (20 launches) so host launch cost is excluded; prints us/launch and TFLOP/s.

    python tools/gemm_bench.py [--one M,N,K,b0,b1,b2,ta,tb]   # --one: run a single shape (for ncu)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from druglamp_b200 import _lib  # noqa: E402

SHAPES = [
    # (M, N, K, batch, ta, tb)
    (16384, 256, 256, (1, 1, 1), 0, 0),
    (16384, 1024, 256, (1, 1, 1), 0, 0),
    (16384, 256, 1024, (1, 1, 1), 0, 0),
    (16384, 2048, 512, (1, 1, 1), 0, 0),
    (16384, 512, 2048, (1, 1, 1), 0, 0),
    (16384, 1024, 256, (1, 1, 1), 0, 1),
    (256, 256, 16384, (1, 1, 1), 1, 1),
    (1024, 256, 16384, (1, 1, 1), 1, 1),
    (512, 2048, 16384, (1, 1, 1), 1, 1),
    (128, 128, 32768, (1, 1, 1), 1, 1),
    (512, 512, 16384, (1, 1, 1), 1, 1),
    (256, 1024, 16384, (1, 1, 1), 1, 1),
    (128, 128, 16384, (1, 1, 1), 1, 1),
    (128, 256, 32768, (1, 1, 1), 1, 1),
    (256, 648, 16384, (1, 1, 1), 1, 1),
    (256, 256, 64, (4, 2, 64), 0, 0),
    (256, 64, 256, (4, 2, 64), 0, 1),
    (256, 512, 128, (1, 1, 64), 0, 0),
    (8192, 8192, 8192, (1, 1, 1), 0, 0),
]


def make(M, N, K, batch, ta, tb, dtype=torch.bfloat16):
    nb = batch[0] * batch[1] * batch[2]
    A = torch.randn((nb, K, M) if ta else (nb, M, K), device="cuda").to(dtype)
    B = torch.randn((nb, K, N) if tb else (nb, N, K), device="cuda").to(dtype)
    # weight-gradient shapes (both operands transposed) produce fp32 like the model does
    C = torch.empty((nb, M, N), device="cuda", dtype=torch.float32 if (ta and tb) else dtype)
    lda, ldb = (M if ta else K), (N if tb else K)
    sa = (M * K, M * K * batch[0], M * K * batch[0] * batch[1])
    sb = (N * K, N * K * batch[0], N * K * batch[0] * batch[1])
    sc = (M * N, M * N * batch[0], M * N * batch[0] * batch[1])

    def run(tile_n=0, split_k=0):
        # weight-gradient shapes accumulate into C like the model does (no memset in the timing)
        _lib.gemm(A, B, C, M=M, N=N, K=K, lda=lda, ldb=ldb, ldc=N, trans_a=bool(ta), trans_b=bool(tb),
                  batch=batch, sa=sa, sb=sb, sc=sc, tile_n=tile_n, split_k=split_k,
                  accumulate=bool(ta and tb and nb == 1))
    return run


def bench(shape, tile_n=0, reps=20, split_k=0):
    run = make(*shape)
    run(tile_n, split_k)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            run(tile_n, split_k)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / (5 * reps)
    M, N, K, batch, _, _ = shape
    fl = 2.0 * M * N * K * batch[0] * batch[1] * batch[2]
    return us, fl / us / 1e6


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        v = [int(x) for x in sys.argv[2].split(",")]
        run = make(v[0], v[1], v[2], (v[3], v[4], v[5]), v[6], v[7])
        for _ in range(3):
            run()
        torch.cuda.synchronize()
    elif len(sys.argv) > 1 and sys.argv[1] == "--splitk":
        # split-K factor sweep on the weight-gradient shapes of the step (K = rows of the batch)
        wg = [(1024, 256, 16384), (256, 1024, 16384), (256, 256, 16384), (768, 256, 16384), (2048, 512, 16384),
              (512, 2048, 16384), (1536, 512, 16384), (512, 512, 16384), (256, 512, 16384), (128, 128, 16384),
              (256, 128, 32768), (256, 392, 32768), (256, 648, 16384), (648, 128, 16384)]
        for (M, N, K) in wg:
            s = (M, N, K, (1, 1, 1), 1, 1)
            line = f"{str((M, N, K)):24s}"
            for tn in (0, 128, 256):
                for sk in (0, 8, 12, 18, 24, 36, 48, 72):
                    try:
                        us, tf = bench(s, tn, split_k=sk)
                        line += f" | bn{tn} sk{sk}: {us:6.1f}"
                    except Exception as e:  # noqa: BLE001
                        line += f" | bn{tn} sk{sk}: ERR"
                line += "\n" + " " * 24
            print(line, flush=True)
    else:
        for s in SHAPES:
            line = f"{str(s):60s}"
            for tn in (0, 64, 128, 256):
                try:
                    us, tf = bench(s, tn)
                    line += f"  bn={tn:3d}: {us:8.1f} us {tf:7.1f} TF/s"
                except Exception as e:  # noqa: BLE001
                    line += f"  bn={tn:3d}: ERR {str(e)[:30]}"
            print(line, flush=True)
