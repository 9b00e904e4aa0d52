"""Offline view of a chrome trace written by tools/graph_profile.py: per stream busy time, the
time no kernel runs at all, and the longest kernels of the last replay in launch order.

    python tools/timeline.py gpurun_out/trace.json [--list] [--alone]

--alone: per kernel name, the time during which it is the ONLY kernel running (a proxy for its share of the
critical path of the multi-stream step), next to its total time.
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
    if "--alone" in sys.argv:
        import re
        ev2 = sorted([(e["ts"], 1, i) for i, e in enumerate(k)] + [(e["ts"] + e["dur"], -1, i) for i, e in enumerate(k)])
        active, last, alone = set(), t0, collections.defaultdict(float)
        for t, d, i in ev2:
            if len(active) == 1:
                alone[next(iter(active))] += t - last
            last = t
            if d == 1:
                active.add(i)
            else:
                active.discard(i)
        by = collections.defaultdict(lambda: [0.0, 0, 0.0])
        for i, e in enumerate(k):
            n = re.sub(r"dl::\(anonymous namespace\)::|dl::<unnamed>::|dl::", "", e["name"].replace("void ", ""))
            b = by[re.sub(r"\(.*", "", n)[:70]]
            b[0] += alone.get(i, 0.0); b[1] += 1; b[2] += e["dur"]
        print(f"# single-kernel time {sum(v[0] for v in by.values()):.1f} us of the span")
        print("alone_us    n  total_us  kernel")
        for n, (a, c, t) in sorted(by.items(), key=lambda kv: -kv[1][0])[:30]:
            print(f"{a:8.1f} {c:4d} {t:9.1f}  {n}")
    if "--list" in sys.argv:
        main_tid = max(per, key=per.get)
        for e in k:
            mark = " " if e["tid"] == main_tid else "*"
            print(f"{e['ts'] - t0:9.1f} {e['dur']:7.1f} {mark}{e['tid']} {e['name'][:110]}")


if __name__ == "__main__":
    main(sys.argv[1])
