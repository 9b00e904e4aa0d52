"""GuidedCrossAttention(128, 1) on its own (SURVEY 8d "config 4" / BASELINE.json configs[3] shape):
query (1200, 256, 128), key = value (S, 256, 128) for S = 290 and S padded to 512, forward + backward
with the raw logit map returned.  Prints pairs/s, the algorithmic TFLOP/s (fwd 4LE^2 + 4SE^2 + 4LSE
per pair, x3 for fwd+bwd) and the HBM rate of the mandatory raw-map store.

    python tools/pgca_bench.py [--dtype bf16|fp32] [--iters 20]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    import druglamp_b200 as D
    from druglamp_b200.modules import GuidedCrossAttention
    dt = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    D.set_compute_dtype(dt)
    L, N, E = 1200, 256, 128
    m = GuidedCrossAttention(E, 1).cuda()
    for S in (290, 512):
        q = torch.randn(L, N, E, device="cuda", dtype=dt).requires_grad_(True)
        k = torch.randn(S, N, E, device="cuda", dtype=dt).requires_grad_(True)
        go = torch.randn(L, N, E, device="cuda", dtype=dt)

        def step():
            q.grad = k.grad = None
            out, raw = m(q, k, k)
            out.backward(go)
            return raw

        for _ in range(5):
            step()
        torch.cuda.synchronize()
        times = []
        for _ in range(a.iters):                  # per-iteration events: eager launches + allocator
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            raw = step()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        times.sort()
        ms = times[len(times) // 2]               # median
        flop = 3 * (4 * L * E * E + 4 * S * E * E + 4 * L * S * E) * N
        print(json.dumps({"workload": f"GuidedCrossAttention(128,1) q(1200,256,128) kv({S},256,128) fwd+bwd",
                          "dtype": a.dtype, "ms": round(ms, 3), "pairs_per_s": round(N / ms * 1e3, 1),
                          "algorithmic_tflops": round(flop / ms * 1e-9, 1),
                          "raw_map_mb": round(raw.numel() * raw.element_size() / 1e6, 1)}))


if __name__ == "__main__":
    main()
