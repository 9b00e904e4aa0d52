"""Per-kernel time of the CAPTURED step (CUDA-graph replay, warm caches, PDL overlap included), from
CUPTI kernel records via torch.profiler -- the in-situ complement of the cold, serialised ncu pass.

    DL_NO_PDL=1 python tools/graph_profile.py [trace.json] > gpurun_out/graph_profile.txt

(with programmatic dependent launch on, a kernel's recorded duration includes its wait for the
previous kernel, so per-kernel numbers are only meaningful with DL_NO_PDL=1)
"""
import collections
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import druglamp_b200 as D  # noqa: E402
from druglamp_b200.models import DrugLAMP  # noqa: E402
from druglamp_b200.synth import make_batch  # noqa: E402
from druglamp_b200.train import StaticBatch, TrainStep  # noqa: E402

REPS = 5


def main():
    dev = torch.device("cuda", 0)
    D.set_compute_dtype(torch.bfloat16)
    torch.manual_seed(1234)
    model = DrugLAMP(384, 640).to(dev)
    model.train()
    model.flatten_parameters()
    ts = TrainStep(model)
    sb = StaticBatch(make_batch(64, seed=1234), dev)
    ts.capture(sb)
    for _ in range(3):
        ts.replay(sb)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(REPS):
            ts.replay(sb)
        torch.cuda.synchronize()
    if len(sys.argv) > 1:                 # chrome trace (stream = tid) for offline timeline analysis
        prof.export_chrome_trace(sys.argv[1])
    agg = collections.defaultdict(lambda: [0, 0.0])
    t_min, t_max = None, None
    for ev in prof.events():
        if ev.device_type is None or "DeviceType.CUDA" not in str(ev.device_type):
            continue
        name = ev.name.replace("(anonymous namespace)::", "").replace("void ", "")
        name = re.sub(r"\(.*", "", name)
        if name.startswith("Memcpy") or name.startswith("Memset"):
            name = name.split(" ")[0]
        agg[name][0] += 1
        agg[name][1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
        tr = ev.time_range
        t_min = tr.start if t_min is None else min(t_min, tr.start)
        t_max = tr.end if t_max is None else max(t_max, tr.end)
    tot = sum(v[1] for v in agg.values())
    span = (t_max - t_min) if t_min is not None else 0.0
    print(f"# captured step, {REPS} replays: sum of kernel durations {tot / REPS / 1000:.3f} ms/step, "
          f"wall span {span / REPS / 1000:.3f} ms/step")
    print("| kernel | launches/step | us/step | share | mean us |")
    print("|---|---:|---:|---:|---:|")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"| `{name[:78]}` | {n / REPS:.1f} | {us / REPS:.1f} | {100 * us / tot:.1f}% | {us / n:.1f} |")


if __name__ == "__main__":
    main()
