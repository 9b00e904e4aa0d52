"""The step's most expensive dl_gemm launches, one launch each between cudaProfilerStart/Stop -- with
their REAL epilogues (bias / GELU / dropout / derivative store / split-K reduction), as recorded from
one eager DrugLAMP step at the bench configuration.

    ncu --profile-from-start off --set full --clock-control none --import-source on \
        -o gpurun_out/gemm_top python tools/gemm_ncu_top.py [--top 8]
    python tools/gemm_ncu_top.py --list        # the order in which they are launched
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import druglamp_b200 as D  # noqa: E402
from druglamp_b200 import _lib as L  # noqa: E402
from druglamp_b200.models import DrugLAMP  # noqa: E402
from druglamp_b200.synth import make_batch  # noqa: E402
from druglamp_b200.train import StaticBatch, TrainStep  # noqa: E402


def describe(rec):
    a = rec["args"]
    epi = []
    if a.bias:
        epi.append("bias")
    if a.act:
        epi.append({1: "gelu", 2: "relu"}.get(a.act, f"act{a.act}"))
    if a.preact_out:
        epi.append("deriv" if a.pre_mode == 1 else "preact")
    if a.mul_aux:
        epi.append("mul_aux")
    if a.drop_p > 0:
        epi.append("dropout")
    if a.residual:
        epi.append("residual")
    if a.accumulate:
        epi.append("accumulate")
    if a.colsum_a:
        epi.append("colsum_a")
    if a.conv_taps:
        epi.append(f"conv{a.conv_taps}")
    if a.kred:
        epi.append(f"kred{a.kred}")
    return f"{rec['shape']} [{'+'.join(epi) or 'plain'}]"


def main():
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 8
    dev = torch.device("cuda", 0)
    D.set_compute_dtype(torch.bfloat16)
    torch.manual_seed(1234)
    model = DrugLAMP(384, 640).to(dev)
    model.train()
    model.flatten_parameters()
    ts = TrainStep(model)
    sb = StaticBatch(make_batch(64, seed=1234), dev)
    for _ in range(2):
        ts._fwd_bwd(sb)
        ts._update()
    torch.cuda.synchronize()
    L.PROFILE = []
    ts._fwd_bwd(sb)
    torch.cuda.synchronize()
    prof, L.PROFILE = L.PROFILE, None
    # time every distinct (shape, epilogue) alone, pick the most expensive in total
    groups = {}
    for r in prof:
        groups.setdefault(describe(r), []).append(r)
    timed = []
    for key, recs in groups.items():
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10):
                L.replay_gemm(recs[0])
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / 30
        timed.append((us * len(recs), us, len(recs), key, recs[0]))
        del g
    timed.sort(key=lambda t: -t[0])
    print("# total_us  n  us/launch  TFLOP/s  shape [epilogue]")
    for tot, us, n, key, rec in timed[:max(top, 40)]:
        print(f"{tot:9.1f}  n={n:3d}  {us:8.1f}  {rec['flops'] / us / 1e6:7.1f}  {key}", flush=True)
    if "--list" in sys.argv:
        return
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _, _, _, _, rec in timed[:top]:
        L.replay_gemm(rec)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
