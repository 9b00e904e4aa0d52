"""dl_layernorm_bwd at the step's shapes: CUDA-graph timing (20 launches per replay), achieved HBM GB/s.
    python tools/ln_bench.py            (DL_LN_CLUSTER=1: column partials reduced over 8-block clusters through DSMEM before the atomics)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from druglamp_b200 import kernels as K  # noqa: E402


def main():
    for rows, C in ((16384, 256), (16384, 512), (16384, 128), (49152, 256)):
        x = torch.randn(rows, C, device="cuda").bfloat16()
        dy = torch.randn(rows, C, device="cuda").bfloat16()
        gamma = torch.randn(C, device="cuda")
        _, mean, rstd = K.layernorm_fwd(x, gamma, gamma, 1e-5)
        dg = torch.zeros(C, device="cuda"); db = torch.zeros(C, device="cuda")
        reps = 20
        K.layernorm_bwd(dy, x, gamma, mean, rstd, acc_into=(dg, db))
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                K.layernorm_bwd(dy, x, gamma, mean, rstd, acc_into=(dg, db))
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / (5 * reps)
        mb = rows * C * 2 * 3 / 1e6
        print(f"layernorm_bwd {rows} x {C}: {us:6.1f} us  {mb / us * 1e3 / 1e3:6.2f} TB/s (3 passes of {mb / 3:.1f} MB, L2-resident working set)")


if __name__ == "__main__":
    main()
