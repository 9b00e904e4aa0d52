"""Offline view of a chrome trace written by tools/graph_profile.py: per stream busy time, the
time no kernel runs at all, and the longest kernels of the last replay in launch order.

    python tools/timeline.py gpurun_out/trace.json [--list]
"""
import collections
import json
import sys


def main(path):
    with open(path) as f:
        ev = json.load(f)["traceEvents"]
    k = [e for e in ev if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in e]
    k.sort(key=lambda e: e["ts"])
    # replays are separated by the largest gaps on the busiest stream; take the last full replay:
    # split on the adamw kernel (last launch of a step)
    ends = [i for i, e in enumerate(k) if "adamw" in e["name"]]
    ends = [i for j, i in enumerate(ends) if j + 1 == len(ends) or ends[j + 1] != i + 1]   # last of each run
    if len(ends) >= 2:
        k = k[ends[-2] + 1:ends[-1] + 1]
    t0, t1 = k[0]["ts"], max(e["ts"] + e["dur"] for e in k)
    print(f"# one replay: {len(k)} launches, span {(t1 - t0):.1f} us")
    per = collections.defaultdict(float)
    for e in k:
        per[e["tid"]] += e["dur"]
    for tid, us in sorted(per.items(), key=lambda kv: -kv[1]):
        print(f"stream {tid}: busy {us:.1f} us ({100 * us / (t1 - t0):.1f}% of the span)")
    # union of all intervals -> idle time; and time with >= 2 kernels in flight
    pts = []
    for e in k:
        pts.append((e["ts"], 1))
        pts.append((e["ts"] + e["dur"], -1))
    pts.sort()
    depth, last, idle, multi = 0, t0, 0.0, 0.0
    for t, d in pts:
        if depth == 0:
            idle += t - last
        elif depth >= 2:
            multi += t - last
        depth += d
        last = t
    print(f"no kernel running: {idle:.1f} us; two or more kernels in flight: {multi:.1f} us")
    if "--list" in sys.argv:
        main_tid = max(per, key=per.get)
        for e in k:
            mark = " " if e["tid"] == main_tid else "*"
            print(f"{e['ts'] - t0:9.1f} {e['dur']:7.1f} {mark}{e['tid']} {e['name'][:110]}")


if __name__ == "__main__":
    main(sys.argv[1])
