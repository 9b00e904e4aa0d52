"""Attribute the PyTorch glue kernels of one DrugLAMP step (copies, adds, fills, embedding backward)
to the source lines that launch them: one eager step under torch.profiler with Python stacks.

    python tools/glue_profile.py > gpurun_out/glue_profile.txt
"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import druglamp_b200 as D  # noqa: E402
from druglamp_b200.models import DrugLAMP  # noqa: E402
from druglamp_b200.synth import make_batch  # noqa: E402
from druglamp_b200.train import StaticBatch, TrainStep  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    D.set_compute_dtype(torch.bfloat16)
    torch.manual_seed(1234)
    model = DrugLAMP(384, 640).to(dev)
    model.train()
    model.flatten_parameters()
    ts = TrainStep(model)
    sb = StaticBatch(make_batch(64, seed=1234), dev)
    for _ in range(2):
        ts._fwd_bwd(sb); ts._update()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, experimental_config=torch._C._profiler._ExperimentalConfig(verbose=True)) as prof:
        ts._fwd_bwd(sb); ts._update()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if not ev.name.startswith("aten::") or ev.self_device_time_total <= 0:
            continue
        # innermost frame inside this repo
        where = "?"
        for fr in ev.stack:
            if "druglamp_b200/" in fr and "train.py" not in fr:
                where = fr.split("druglamp_b200/")[-1]
                break
        if where == "?" and ev.stack:
            where = "stack0: " + ev.stack[0][-50:]
        shapes = ""
        key = (ev.name, where, shapes)
        agg[key][0] += 1
        agg[key][1] += ev.self_device_time_total
    tot = sum(v[1] for v in agg.values())
    print(f"aten ops with device time: {tot:.0f} us total (eager, not a bench number)")
    for (name, where, shapes), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
        print(f"{us:8.1f} us  n={n:3d}  {name:34s} {where:60s} {shapes}")


if __name__ == "__main__":
    main()
