"""Summarise an `ncu --csv` metrics log of one DrugLAMP step (bench.py --ncu-step) per kernel:
launches, total / mean duration, share of the step, DRAM bytes, tensor-pipe activity.

    python tools/ncu_summary.py gpurun_out/step_metrics.csv > profiles/rNN_step_summary.md
"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    per = collections.OrderedDict()
    for r in rows:
        key = r["ID"]
        d = per.setdefault(key, {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")})
        v = float(r["Metric Value"].replace(",", "")) if r["Metric Value"] not in ("", "n/a") else 0.0
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        if m.startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[m] = v
    agg = collections.defaultdict(lambda: collections.defaultdict(float))
    for d in per.values():
        a = agg[d["name"]]
        a["n"] += 1
        a["us"] += d.get("gpu__time_duration.sum", 0.0)
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
        a["tensor_x_us"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * \
            d.get("gpu__time_duration.sum", 0.0)
    total = sum(a["us"] for a in agg.values())
    print(f"# One DrugLAMP step under ncu ({len(per)} kernel launches, {total / 1000:.2f} ms summed kernel time; "
          "cold-cache, serialised: compare SHARES, not absolutes)\n")
    print("| kernel | launches | total us | share | mean us | DRAM read MB | DRAM write MB | tensor pipe % (time-weighted) |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        tp = a["tensor_x_us"] / a["us"] if a["us"] else 0.0
        print(f"| `{name[:70]}` | {int(a['n'])} | {a['us']:.1f} | {100 * a['us'] / total:.1f}% | {a['us'] / a['n']:.1f} | "
              f"{a['rd'] / 1e6:.1f} | {a['wr'] / 1e6:.1f} | {tp:.1f} |")
    mine = sum(a["us"] for n, a in agg.items() if n.startswith("dl::"))
    print(f"\nHand-written dl::* kernels: {100 * mine / total:.1f}% of the summed kernel time; "
          f"the rest is PyTorch glue (copies, adds, embedding, index ops).")
    g = [a for n, a in agg.items() if "gemm_tc_kernel" in n]
    if g:
        n = sum(a["n"] for a in g)
        print(f"\ngemm_tc_kernel: {int(n)} launches, DRAM traffic per launch (mean) "
              f"{(sum(a['rd'] + a['wr'] for a in g)) / n / 1e6:.2f} MB.")


if __name__ == "__main__":
    main(sys.argv[1])
