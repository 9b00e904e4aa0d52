#!/usr/bin/env python
"""DrugLAMP hot-path benchmark: DTI pairs/sec, forward + backward (+ AdamW), on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload = BASELINE.json configs[1]: full DrugLAMP (GCN drug + CNN protein + LLM adaptors + PGCA +
MHLA + PMMA + MLP head), BioSNAP-shaped synthetic pairs, batch 64 per GPU, bf16, train mode
(PMMA dropout 0.1 live).  Data-parallel over pairs (weak scaling): every rank steps its own 64
pairs, gradients are all-reduced once per step over NCCL.

One JSON line on rank 0:
  value     whole-job pairs/s with inputs resident in HBM (CUDA-graph replay, CUDA events, max over ranks)
  e2e       same metric through the public API with HOST inputs: pinned H2D of every input + D2H of the
            loss every step; the LLM embeddings cross PCIe as the dataset yields them (packed rows,
            druglamp_b200/collate.py) and are padded / tiled on the device like the reference collate
  e2e_dense the same with the reference collate's dense fp32 tensors crossing PCIe (6.9 MB/pair)
  roofline  the dominant kernel (dl_gemm's tcgen05 kernel): algorithmic FLOPs / CUDA-event time of
            every launch in one instrumented step, against the measured bf16 peak
  cpu_baseline  the oracle port of the reference's CPU path timed on this box's host cores (N=1 only)
`--impl reference` times that CPU path as its own arm (same metric / config / unit).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "dti_pairs_per_sec_fwd_bwd"
UNIT = "pairs/s"
BATCH = 64
FLOP_PER_PAIR_FWD_BWD = 24.61e9      # GEMM/bmm/conv 2MNK, dense padded shapes, fwd+bwd
N_DISTINCT_BATCHES = 3


def config(n_gpus):
    return {"workload": "DrugLAMP full (BASELINE.json configs[1]): GCN+CNN+LLM adaptors+PGCA+MHLA+PMMA+MLP, "
                        "BioSNAP-shaped synthetic pairs, train mode (dropout 0.1), fwd+BCE+bwd+AdamW",
            "batch_per_gpu": BATCH, "global_batch": BATCH * n_gpus, "parallelism": f"dp{n_gpus}",
            "protein_tokens": 2304, "drug_nodes": 512, "llm_dims": [640, 384],
            "l2_policy": f"inputs larger than L2: {N_DISTINCT_BATCHES} distinct resident batches of ~440 MB rotate, "
                         "each step reads one in full"}


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p.update(json.load(f))
            p["source"] = "measured"
    except Exception:
        pass
    return p


# ------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------- CPU path
def cpu_port_step_fn(batch_size, kind="DrugLAMP", seed=4321):
    """The oracle port of the reference's CPU path: zero_grad -> forward -> BCE -> backward -> AdamW."""
    from oracle import restatement as R
    from druglamp_b200.synth import make_batch
    with open(os.path.join(ROOT, "tests", "golden", "state_shapes.json")) as f:
        shapes = {k: tuple(v) for k, v in json.load(f).items()}
    sd = R.deterministic_state(shapes)
    params = [v.requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    opt = torch.optim.AdamW(params, lr=1e-4)
    b = make_batch(batch_size, seed=seed)

    def step():
        opt.zero_grad(set_to_none=True)
        o = R.druglamp_forward(sd, kind, b.graph.src, b.graph.dst, b.graph.ndata["h"], batch_size,
                               b.vp, b.xd, b.xp, True)
        _, loss = R.binary_cross_entropy(o["score"], b.y)
        loss.backward()
        opt.step()
        return float(loss.item())
    return step


def cpu_reference_step_fn(batch_size, kind="DrugLAMP", seed=4321):
    """The reference's OWN modules (model/DrugLAMP.py built by DrugLAMPBase.__init__, unmodified:
    /root/reference in the build container, its byte-compiled copy oracle/_ref on the GPU box) stepped
    as trainer.py:196-229 steps them on non-SSL/CM epochs: forward -> binary_cross_entropy ->
    zero_grad -> backward -> AdamW.step.  DGL's update_all is the index_add_ restatement of ref_shim."""
    from oracle import ref_shim
    from druglamp_b200.synth import make_batch
    m = ref_shim.build_reference_model(kind)
    from model.basic_model import binary_cross_entropy
    m.train()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4)
    b = make_batch(batch_size, seed=seed)

    def step():
        g = ref_shim.FakeGraph(b.graph.src, b.graph.dst, b.graph.num_nodes(), batch_size, b.graph.ndata["h"])
        out = m(g, b.vp, b.xd, b.xp)
        opt.zero_grad(set_to_none=True)
        _, loss = binary_cross_entropy(out[4], b.y)
        loss.backward()
        opt.step()
        return float(loss.item())
    return step


def time_cpu(batch_size, steps, warmup):
    """-> (pairs/s from the median step, kind): the reference's modules when they are available
    (kind "reference"), else the oracle port of them (kind "port")."""
    torch.set_num_threads(os.cpu_count() or 1)
    kind = "port"
    try:
        from oracle import ref_shim
        if ref_shim.available():
            step, kind = cpu_reference_step_fn(batch_size), "reference"
    except Exception as e:                                   # pragma: no cover
        print(f"bench: reference modules unavailable ({e}); timing the oracle port", file=sys.stderr)
    if kind == "port":
        step = cpu_port_step_fn(batch_size)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return batch_size / statistics.median(ts), kind


CPU_MAX_STEPS = 12      # a 64-pair CPU step takes ~1.5 s on 16 cores: bound the arm to ~20 s


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    bs = BATCH
    warm = min(args.warmup, 1)
    steps = max(1, min(args.steps, CPU_MAX_STEPS))
    pps, kind = time_cpu(bs, steps, warm)
    cores = os.cpu_count() or 1
    what = ("the reference's own modules (model/DrugLAMP.py, unmodified)" if kind == "reference"
            else "oracle port of the reference modules")
    line = {"impl": "reference", "metric": METRIC, "value": pps, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": 1000.0 * bs / pps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config(args.gpus),
            "cpu_baseline": {"value": pps, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{steps} timed steps of {bs} pairs after {warm} warm-up (the --steps request "
                                       f"is capped at {CPU_MAX_STEPS}; forward+BCE+zero_grad+backward+AdamW), "
                                       f"{what}, torch CPU fp32, all host threads, median step"},
            "e2e": {"value": pps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def _teardown(world, *drop):
    """Leave cleanly at N > 1.  The captured step graphs contain NCCL kernels; a communicator that is
    destroyed while such graphs are alive (or while another rank is still busy) can block forever, and
    a bench that printed its line but never exits is a hung run for the driver.  So: graphs first, then
    a barrier, then destroy -- and a watchdog that ends the process if the teardown does not return."""
    if world <= 1:
        return
    import gc
    import threading
    sys.stdout.flush()
    threading.Timer(45.0, lambda: os._exit(0)).start()
    for d in drop:
        try:
            d.clear()
        except Exception:
            pass
    gc.collect()
    torch.cuda.synchronize()
    try:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    finally:
        sys.stdout.flush()
        os._exit(0)


# ------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import druglamp_b200 as D
    from druglamp_b200 import _lib as L
    from druglamp_b200.models import DrugLAMP
    from druglamp_b200.synth import make_batch
    from druglamp_b200.train import StaticBatch, TrainStep

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
    D.set_compute_dtype(torch.bfloat16)
    L.lib()

    torch.manual_seed(1234)
    model = DrugLAMP(384, 640).to(dev)
    model.train()
    model.flatten_parameters()
    ts = TrainStep(model, world_size=world)
    if world > 1:                      # replicas start from rank 0's weights
        torch.distributed.broadcast(ts.flat.flat, 0)

    raw = [make_batch(BATCH, seed=1234 + rank * 100 + i) for i in range(N_DISTINCT_BATCHES)]
    batches = [StaticBatch(b, dev) for b in raw]
    hosts = [b.host_copy(pin=True) for b in batches]
    h2d = sum(t.numel() * t.element_size() for t in hosts[0])
    # packed wire format (druglamp_b200/collate.py): the embedding rows as the dataset yields them,
    # before the reference collate tiles / pads them on the host
    hosts_packed = [sb.host_copy_packed(b, pin=True) for sb, b in zip(batches, raw)]
    if args.ncu_step:
        # profiling aid (never a bench value): one eager step between cudaProfilerStart/Stop so that
        #   ncu --profile-from-start off ... python bench.py --ncu-step
        # sees exactly the kernels of one step (forward, backward on the autograd thread, optimizer)
        for _ in range(2):
            ts._fwd_bwd(batches[0]); ts._update()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        ts._fwd_bwd(batches[0]); ts._update()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    for b in batches:
        ts.capture(b)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM, CUDA-graph replay ------------------------------------
    for i in range(max(3, args.warmup)):
        ts.replay(batches[i % len(batches)])
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        ts.replay(batches[i % len(batches)])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # a 20-step region lasts ~0.1 s (one nvidia-smi sample): when K is small, time 200 more steps the
    # same way so the clock record covers a real stretch of load; `value` stays the exact-K figure
    long_ms, long_steps = None, 200
    if args.steps < long_steps:
        barrier()
        e0.record()
        for i in range(long_steps):
            ts.replay(batches[i % len(batches)])
        e1.record()
        barrier()
        long_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    loss_value = float(ts.loss.item())

    # ---- e2e: host inputs -> pinned H2D -> step -> D2H of the loss, every step -----------------
    # Every step's inputs cross PCIe inside the timed region; the copy of step i+1 runs on a copy
    # stream while step i computes (buffers rotate, events order reuse), as a real input pipeline does.
    e2e_steps = max(3, min(args.steps, 60))
    nb = len(batches)
    copy_stream = torch.cuda.Stream()
    copied = [torch.cuda.Event() for _ in range(nb)]
    consumed = [torch.cuda.Event() for _ in range(nb)]
    main = torch.cuda.current_stream()

    h2d_bytes = {"dense": 0, "packed": 0}

    def enqueue_copy(i, fmt):
        j = i % nb
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[j])      # the step that last read these buffers is done
            if fmt == "packed":                      # untiled rows over PCIe, tiled on the device
                h2d_bytes[fmt] = batches[j].load_from_packed(hosts_packed[j])
            else:                                    # the reference collate's dense tensors
                h2d_bytes[fmt] = batches[j].load_from(hosts[j])
            copied[j].record(copy_stream)

    def e2e_loop(n, fmt):
        for j in range(nb):
            consumed[j].record(main)
        enqueue_copy(0, fmt)
        for i in range(n):
            j = i % nb
            main.wait_event(copied[j])
            out = ts.replay(batches[j])
            consumed[j].record(main)
            if i + 1 < n:
                enqueue_copy(i + 1, fmt)
            out.item()                               # D2H read of the step's loss (4 bytes), every step

    e2e_times = {}
    for fmt in ("dense", "packed"):
        e2e_loop(3, fmt)
        barrier()
        t0 = time.perf_counter()
        e2e_loop(e2e_steps, fmt)
        barrier()
        e2e_times[fmt] = time.perf_counter() - t0

    t = torch.tensor([ms, e2e_times["dense"], e2e_times["packed"], long_ms or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, e2e_dense_s, e2e_s = float(t[0]), float(t[1]), float(t[2])
    long_ms = float(t[3]) if long_ms is not None else None

    # ---- roofline of the dominant kernel: every dl_gemm launch of one eager step, CUDA events ---
    roof = None
    if rank == 0:
        pk = peaks()

        # only rank 0 runs this pass, so it must not contain a collective: the overlapped gradient
        # all-reduce lives inside _fwd_bwd and is switched off here
        ts.overlap = False
        model._pmma_grads_ready = None

        def local_step():
            ts._fwd_bwd(batches[0])
            ts._update()
        local_step()
        torch.cuda.synchronize()
        L.PROFILE = []
        local_step()
        torch.cuda.synchronize()
        prof, L.PROFILE = L.PROFILE, None
        # Re-issue exactly those launches back to back inside one CUDA graph and time the replays
        # with CUDA events on the launching stream: kernel time without host launch gaps.
        gg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gg):
            for rec in prof:
                L.replay_gemm(rec)
        gg.replay()
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        g0.record()
        for _ in range(reps):
            gg.replay()
        g1.record()
        torch.cuda.synchronize()
        gemm_ms = g0.elapsed_time(g1) / reps
        flops = sum(r["flops"] for r in prof)
        if os.environ.get("DL_BENCH_DUMP"):
            agg = {}
            for r in prof:
                d = agg.setdefault(str(r["shape"]), {"n": 0, "flops": 0.0, "rec": r})
                d["n"] += 1; d["flops"] += r["flops"]
            for d in agg.values():                 # time each distinct shape alone (10 launches per graph)
                g1 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1):
                    for _ in range(10):
                        L.replay_gemm(d["rec"])
                g1.replay()
                torch.cuda.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(3):
                    g1.replay()
                a1.record()
                torch.cuda.synchronize()
                d["us"] = a0.elapsed_time(a1) * 1000 / 30
                del g1
            os.makedirs(os.path.dirname(os.environ["DL_BENCH_DUMP"]) or ".", exist_ok=True)
            with open(os.environ["DL_BENCH_DUMP"], "w") as f:
                f.write(f"# {len(prof)} tensor-core launches per step (dl_gemm + fused dl_ffn_fwd / dl_ffn_bwd), {flops / 1e9:.1f} GFLOP, {gemm_ms:.3f} ms back-to-back\n")
                f.write("# total_us  n  us/launch  TFLOP/s  (M,N,K,batch,ta,tb)\n")
                for k, d in sorted(agg.items(), key=lambda kv: -kv[1]["n"] * kv[1]["us"]):
                    f.write(f"{d['n'] * d['us']:9.1f}  n={d['n']:3d}  {d['us']:8.1f}  {d['flops'] / d['n'] / d['us'] / 1e6:7.1f}  {k}\n")
        del gg
        achieved = flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        peak = pk["bf16_tflops_sustained"]
        step_ms = ms / args.steps
        traffic, traffic_src = None, None
        tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "gemm_traffic.json")
        if os.path.exists(tpath):              # committed ncu measurement (never taken inside a bench run)
            with open(tpath) as f:
                tj = json.load(f)
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
        roof = {"kernel": "gemm_tc_kernel (dl_gemm: TMA + tcgen05.mma, CTA pairs per shape, bf16 in / fp32 TMEM accumulate) "
                          "+ ffn_chain_kernel (dl_ffn_fwd / dl_ffn_bwd: the two FFN GEMMs chained on chip)",
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_unit": "DRAM bytes per launch (mean)",
                "traffic_source": traffic_src, "peak_source": pk["source"] + " bf16_tflops_sustained",
                "launches_per_step": len(prof), "flops_per_launch_avg": flops / max(len(prof), 1),
                "gemm_flops_per_step": flops, "avg_launch_us": 1000.0 * gemm_ms / max(len(prof), 1),
                "gemm_ms_per_step": gemm_ms,
                "how": "all dl_gemm / dl_ffn launches of one step re-issued back to back in a CUDA graph, CUDA events, mean of 5 replays",
                "share_of_step": min(1.0, gemm_ms / step_ms) if step_ms > 0 else None,
                "whole_step_tflops": FLOP_PER_PAIR_FWD_BWD * BATCH / (step_ms * 1e-3) / 1e12,
                "whole_step_frac": FLOP_PER_PAIR_FWD_BWD * BATCH / (step_ms * 1e-3) / 1e12 / peak}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        pps, kind = time_cpu(BATCH, 5, 1)
        cpu = {"value": pps, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": kind,
               "sample": f"5 timed steps of {BATCH} pairs after 1 warm-up (fwd+BCE+zero_grad+bwd+AdamW, median), "
                         + ("the reference's own modules" if kind == "reference" else "oracle port of the reference modules")
                         + " on torch CPU fp32 with all host threads"}

    if rank == 0:
        pairs = BATCH * world
        line = {"metric": METRIC, "value": pairs * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic", "config": config(world), "clocks": clocks,
                # e2e: the reference API's own host tensors (utils.multimodality_collate_func: dense fp32
                # embeddings, tokens, node features, the batched graph's raw edge list) -> pinned H2D ->
                # device-side CSR construction (dl_csr_build) -> step -> loss D2H, every step.
                # e2e_packed: the embeddings cross PCIe as the DATASET yields them (per-sample rows,
                # druglamp_b200.collate.pack_rows) and are padded / tiled on the device (dl_expand_rows,
                # bit-identical to the reference collate's tensors).
                "e2e": {"value": pairs * e2e_steps / e2e_dense_s, "unit": UNIT,
                        "h2d_bytes_per_step": h2d_bytes["dense"], "d2h_bytes_per_step": 4, "steps": e2e_steps,
                        "wire_format": "dense fp32 tensors of utils.multimodality_collate_func + raw (src, dst) "
                                       "edge list; PCIe-bound: 6.9 MB per pair"},
                "e2e_packed": {"value": pairs * e2e_steps / e2e_s, "unit": UNIT,
                               "h2d_bytes_per_step": h2d_bytes["packed"], "d2h_bytes_per_step": 4,
                               "steps": e2e_steps,
                               "wire_format": "packed per-sample embedding rows, padded/tiled on the device"},
                "gpu_launches": ts.launches_per_step * args.steps, "gpu_launches_per_step": ts.launches_per_step,
                "roofline": roof, "cpu_baseline": cpu, "loss": loss_value}
        if long_ms is not None:
            line["long_run"] = {"steps": long_steps, "ms_per_step": long_ms / long_steps,
                                "value": pairs * long_steps / (long_ms * 1e-3), "unit": UNIT}
        print(json.dumps(line), flush=True)
    _teardown(world, ts._graphs)


# ------------------------------------------------------------------------------------- other configs
def _dist_setup():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
    return rank, local, world, dev


def _timed(fn, steps, warmup, world):
    """warm up, then time `steps` calls with CUDA events between barriers; max over ranks -> ms total"""
    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
    for _ in range(max(3, warmup)):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t[0])


def synthetic_global_meta(n, seed=99):
    """BindingDB-shaped ids for a global batch: ~5 drugs per protein, 45 % positives."""
    import numpy as np
    rng = np.random.default_rng(seed)
    n_prot, n_drug = max(1, n // 5), max(1, int(n * 0.8))
    return [{"Prot_ID": f"P{int(rng.integers(n_prot))}", "Drug_ID": f"D{int(rng.integers(n_drug))}",
             "Y": int(rng.random() < 0.45)} for _ in range(n)]


def run_2c2p(args):
    """BASELINE.json configs[2]: DrugLAMP2C2P, classification + contrastive (2C2P) loss every step,
    global batch 4096 = 4096 / G pairs per GPU stepped as 64-pair micro-batches with gradient
    accumulation, NCCL all-gathered pooled features as global negatives (druglamp_b200/contrastive.py)."""
    import druglamp_b200 as D
    from druglamp_b200 import _lib as L
    from druglamp_b200.contrastive import ContrastiveStep
    from druglamp_b200.models import DrugLAMP2C2P
    from druglamp_b200.modules import CrossModality
    from druglamp_b200.synth import make_batch
    from druglamp_b200.train import StaticBatch
    rank, local, world, dev = _dist_setup()
    D.set_compute_dtype(torch.bfloat16)
    L.lib()
    GLOBAL = args.global_batch
    n_local = GLOBAL // world
    n_micro = n_local // BATCH
    assert n_micro * BATCH * world == GLOBAL, "global batch must be a multiple of 64 * GPUs"
    torch.manual_seed(1234)
    model = DrugLAMP2C2P(384, 640).to(dev)
    model.train()
    model.flatten_parameters()
    targets = CrossModality.prepare(synthetic_global_meta(GLOBAL))
    cs = ContrastiveStep(model, targets, n_local, world_size=world, rank=rank)
    if world > 1:
        torch.distributed.broadcast(cs.flat.flat, 0)
    distinct = [StaticBatch(make_batch(BATCH, seed=1234 + rank * 100 + i), dev) for i in range(N_DISTINCT_BATCHES)]
    cs.capture(distinct, n_micro)
    micro = [distinct[i % len(distinct)] for i in range(n_micro)]
    steps = args.steps if args.steps is not None else 8
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = _timed(lambda: cs.step(micro, graphs=True), steps, min(args.warmup, 3), world)
    clocks = sampler.stop() if rank == 0 else None
    loss = float(cs.cls_loss + cs.cm_loss)
    # share of the contrastive part (latents + P x D similarity + triplet loss, forward + backward)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        cs._g_cm.replay()
    e1.record()
    torch.cuda.synchronize()
    cm_ms = e0.elapsed_time(e1) / 5
    if rank == 0:
        step_ms = ms / steps
        line = {"metric": METRIC, "value": GLOBAL * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": steps, "warmup": max(3, min(args.warmup, 3)), "ms_per_step": step_ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": "DrugLAMP2C2P contrastive pretraining (BASELINE.json configs[2]): classification + "
                                       "2C2P triplet loss with global in-batch negatives every step, BindingDB-shaped "
                                       "synthetic ids (~5 drugs per protein), fwd+bwd+AdamW",
                           "global_batch": GLOBAL, "batch_per_gpu": n_local, "micro_batch": BATCH,
                           "micro_batches_per_gpu": n_micro, "parallelism": f"dp{world}",
                           "unique_proteins": int(targets.G.shape[0]), "unique_drugs": int(targets.G.shape[1]),
                           "l2_policy": f"{N_DISTINCT_BATCHES} distinct resident 64-pair input batches (~440 MB each) rotate "
                                        "through the micro-steps; the 4096 ids / labels are all distinct entries"},
                "clocks": clocks,
                "contrastive": {"ms_per_step": cm_ms, "share_of_step": cm_ms / step_ms,
                                "similarity_gflop": 2.0 * targets.G.shape[0] * targets.G.shape[1] * 256 / 1e9,
                                "gathered_bytes_per_rank": 4 * GLOBAL * 128 * 4},
                "e2e": None, "gpu_launches": cs.launches * steps, "gpu_launches_per_step": cs.launches,
                "loss": loss}
        print(json.dumps(line), flush=True)
    _teardown(world, cs._g_feat, cs._g_back)


def run_pgca(args):
    """BASELINE.json configs[3]: long-sequence stress of the co-attention alone -- GuidedCrossAttention(128, 1),
    protein 1200 residues x drug 290 atoms, batch 256, forward + backward, raw logit map returned."""
    import druglamp_b200 as D
    from druglamp_b200 import _lib as L
    from druglamp_b200.modules import GuidedCrossAttention
    rank, local, world, dev = _dist_setup()
    D.set_compute_dtype(torch.bfloat16)
    Lq, S, N, E = 1200, 290, 256, 128
    torch.manual_seed(5 + rank)
    m = GuidedCrossAttention(E, 1).to(dev)
    q = torch.randn(Lq, N, E, device=dev, dtype=torch.bfloat16).requires_grad_(True)
    k = torch.randn(S, N, E, device=dev, dtype=torch.bfloat16).requires_grad_(True)
    go = torch.randn(Lq, N, E, device=dev, dtype=torch.bfloat16)

    def step():
        q.grad = k.grad = None
        for p_ in m.parameters():
            p_.grad = None
        out, raw = m(q, k, k)
        out.backward(go)
        return raw
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    n0 = L.launch_count()
    with torch.cuda.graph(g):
        raw = step()
    launches = L.launch_count() - n0
    steps = args.steps if args.steps is not None else 100
    ms = _timed(g.replay, steps, args.warmup, world)
    if rank == 0:
        flop = 3 * (4 * Lq * E * E + 4 * S * E * E + 4 * Lq * S * E) * N
        step_ms = ms / steps
        print(json.dumps({"metric": METRIC, "value": N * world * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
                          "steps": steps, "warmup": max(3, args.warmup), "ms_per_step": step_ms,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                          "data": "synthetic",
                          "config": {"workload": "GuidedCrossAttention(128, 1) alone (BASELINE.json configs[3]): query "
                                                 "(1200, 256, 128), key = value (290, 256, 128), fwd+bwd, raw map returned",
                                     "batch_per_gpu": N, "parallelism": f"replicas x{world}",
                                     "l2_policy": "working set 0.5 GB per step (> L2)"},
                          "algorithmic_tflops": flop / step_ms * 1e-9,
                          "raw_map_bytes": raw.numel() * raw.element_size(), "e2e": None,
                          "gpu_launches": launches * steps, "gpu_launches_per_step": launches}), flush=True)
    _teardown(world)


def run_infer(args):
    """BASELINE.json configs[4]: forward-only scoring sweep, eval mode, bf16, 128 pairs per GPU
    (batch 1024 on 8 GPUs), data-parallel replicas with no exchange step."""
    import druglamp_b200 as D
    from druglamp_b200.infer import InferStep
    from druglamp_b200.models import DrugLAMP
    from druglamp_b200.synth import make_batch
    from druglamp_b200.train import StaticBatch
    rank, local, world, dev = _dist_setup()
    D.set_compute_dtype(torch.bfloat16)
    torch.manual_seed(1234)
    model = DrugLAMP(384, 640).to(dev)
    st = InferStep(model)
    B = 128
    batches = [StaticBatch(make_batch(B, seed=77 + rank * 10 + i), dev) for i in range(3)]
    for b in batches:
        st.capture(b)
    it = [0]

    def step():
        st.replay(batches[it[0] % 3])
        it[0] += 1
    steps = args.steps if args.steps is not None else 100
    ms = _timed(step, steps, args.warmup, world)
    if rank == 0:
        step_ms = ms / steps
        print(json.dumps({"metric": "dti_pairs_per_sec_fwd", "value": B * world * steps / (ms * 1e-3), "unit": UNIT,
                          "n_gpus": world, "steps": steps, "warmup": max(3, args.warmup), "ms_per_step": step_ms,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                          "data": "synthetic",
                          "config": {"workload": "DrugLAMP eval-mode forward (BASELINE.json configs[4]): kinase-sweep-shaped "
                                                 "scoring, BatchNorm on running statistics, no dropout",
                                     "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"replicas x{world}",
                                     "l2_policy": "3 distinct resident batches of ~0.9 GB rotate"},
                          "algorithmic_tflops_per_gpu": 8.277 * B / step_ms, "e2e": None,
                          "gpu_launches": st.launches_per_step * steps,
                          "gpu_launches_per_step": st.launches_per_step}), flush=True)
    _teardown(world)


def run_trainer(args):
    """The reference's training step with its auxiliary losses (trainer.py:179-231):
    DrugLAMP2C2P, 64 pairs, bf16, three AdamW optimisers over all parameters, stepped eagerly (the MLM mask
    is sampled per step and the 2C2P label matrix is built from the batch's meta dicts on the host, as the
    reference does).  `value` = an SSL + 2C2P epoch's step; the other epoch kinds are listed beside it."""
    import druglamp_b200 as D
    from druglamp_b200 import _lib as L
    from druglamp_b200.models import DrugLAMP2C2P
    from druglamp_b200.synth import make_batch
    from druglamp_b200.train import StaticBatch
    from druglamp_b200.trainer_step import TrainerStep
    rank, local, world, dev = _dist_setup()
    D.set_compute_dtype(torch.bfloat16)
    torch.manual_seed(1234)
    model = DrugLAMP2C2P(384, 640).to(dev)
    model.train()
    model.flatten_parameters()
    raw = [make_batch(BATCH, seed=1234 + rank * 100 + i) for i in range(N_DISTINCT_BATCHES)]
    batches = [StaticBatch(b, dev) for b in raw]
    steps = args.steps if args.steps is not None else 30
    out = {}
    launches = 0
    for name, (ssl, cm), wiped in (("cls", (False, False), False), ("cls+ssl", (True, False), False),
                                   ("cls+ssl+2c2p", (True, True), False),
                                   ("cls+ssl+2c2p, wiped backward passes run too", (True, True), True)):
        ts = TrainerStep(model, run_wiped_backward=wiped)
        it = [0]

        def step():
            i = it[0] % len(batches)
            ts.step(batches[i], meta=raw[i].meta, compute_ssl=ssl, compute_cm=cm)
            it[0] += 1
        n0 = L.launch_count()
        step()
        launches = L.launch_count() - n0
        ms = _timed(step, steps, args.warmup, world)
        out[name] = {"ms_per_step": ms / steps, "pairs_per_s": BATCH * world * steps / (ms * 1e-3),
                     "gpu_launches_per_step": launches}
    if rank == 0:
        head = out["cls+ssl+2c2p"]
        print(json.dumps({"metric": METRIC, "value": head["pairs_per_s"], "unit": UNIT, "n_gpus": world,
                          "steps": steps, "warmup": max(3, args.warmup), "ms_per_step": head["ms_per_step"],
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                          "data": "synthetic",
                          "config": {"workload": "reference training step with auxiliary losses (trainer.py:179-231): "
                                                 "DrugLAMP2C2P forward, classification + SSL x0.1 + 2C2P losses, three AdamW "
                                                 "over all parameters; eager launches (host-side mask sampling / label matrix), "
                                                 "no gradient all-reduce in this line",
                                     "batch_per_gpu": BATCH, "parallelism": f"replicas x{world}",
                                     "l2_policy": f"{N_DISTINCT_BATCHES} distinct resident batches of ~440 MB rotate"},
                          "epoch_kinds": out, "e2e": None, "gpu_launches": head["gpu_launches_per_step"] * steps,
                          "gpu_launches_per_step": head["gpu_launches_per_step"]}), flush=True)
    _teardown(world)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--config", default="druglamp", choices=["druglamp", "2c2p", "pgca", "infer", "trainer"],
                    help="druglamp = BASELINE configs[1] (the headline, default); 2c2p = configs[2]; "
                         "pgca = configs[3]; infer = configs[4]; trainer = the reference training step with its "
                         "SSL / 2C2P losses and three optimisers (trainer.py:179-231)")
    ap.add_argument("--global-batch", type=int, default=4096, help="--config 2c2p only")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ncu-step", action="store_true", help="run one eager step in an NVTX range and exit (for ncu)")
    args = ap.parse_args()
    if args.config != "druglamp" and args.impl == "b200":
        {"2c2p": run_2c2p, "pgca": run_pgca, "infer": run_infer, "trainer": run_trainer}[args.config](args)
        return
    if args.steps is None:
        args.steps = 200
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
