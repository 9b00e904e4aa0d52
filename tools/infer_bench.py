"""Forward-only scoring throughput (BASELINE.json configs[4] shape: eval mode, bf16, 128 pairs per GPU,
data-parallel replicas without any exchange step; `bench.py` keeps timing configs[1]).

    python tools/infer_bench.py [--batch 128] [--steps 30] [--warmup 5]

Three distinct synthetic batches (~0.9 GB each, larger than L2) rotate through CUDA-graph replays of
`druglamp_b200.infer.InferStep`; timing with CUDA events on the launching stream.  Prints one JSON line:
pairs/s, ms per batch, kernel launches per batch and the algorithmic TFLOP/s (8.277 GFLOP forward per
pair)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    a = ap.parse_args()
    import druglamp_b200 as D
    from druglamp_b200.infer import InferStep
    from druglamp_b200.models import DrugLAMP
    from druglamp_b200.synth import make_batch
    from druglamp_b200.train import StaticBatch
    D.set_compute_dtype(torch.bfloat16)
    dev = torch.device("cuda", 0)
    torch.manual_seed(1234)
    model = DrugLAMP(384, 640).to(dev)
    st = InferStep(model)
    batches = [StaticBatch(make_batch(a.batch, seed=77 + i), dev) for i in range(3)]
    for b in batches:
        st.capture(b)
    for i in range(max(3, a.warmup)):
        st.replay(batches[i % 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        n, loss = st.replay(batches[i % 3])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    assert bool(torch.isfinite(n).all()) and bool(torch.isfinite(loss))
    print(json.dumps({"workload": f"DrugLAMP eval forward, batch {a.batch}, bf16, 1 GPU (configs[4] shape)",
                      "pairs_per_s": round(a.batch / ms * 1e3, 1), "ms_per_batch": round(ms, 3),
                      "gpu_launches": st.launches_per_step,
                      "algorithmic_tflops": round(8.277 * a.batch / ms, 1), "steps": a.steps}))


if __name__ == "__main__":
    main()
