"""N-GPU check of the 2C2P loss with NCCL all-gathered (global) negatives on the real kernels:
every rank's loss equals the single-process loss on the rank-ordered concatenated batch and the
gradient reaching its local features equals that rank's slice of the full-batch gradient (scaled
by the world size: averaged-DDP convention).  The CPU twin of this check (gloo, oracle arithmetic)
is tests/test_parallel_gloo_cpu.py.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/nccl_cm_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    import druglamp_b200 as D
    from druglamp_b200.modules import CrossModality
    from druglamp_b200.parallel import global_cross_modality_loss
    from druglamp_b200.synth import make_batch
    D.set_compute_dtype(torch.float32)
    torch.manual_seed(0)
    cm = CrossModality(hidden_size=128).to(dev)
    cm.train(True)
    B = 12
    g = torch.Generator().manual_seed(21)
    meta = make_batch(B * world, seed=21, drugs_per_protein=3.0).meta
    feats = [torch.randn(B * world, 6, 128, generator=g).to(dev) for _ in range(4)]
    loc = [f[rank * B:(rank + 1) * B].clone().requires_grad_(True) for f in feats]
    loss = global_cross_modality_loss(cm, *loc, meta[rank * B:(rank + 1) * B], pool_fn=lambda s: s.mean(1))
    loss.backward()
    # single-process reference on the concatenated batch (same module, same kernels, no collective)
    cm2 = CrossModality(hidden_size=128).to(dev)
    cm2.load_state_dict(cm.state_dict())
    cm2.train(True)
    full = [f.clone().requires_grad_(True) for f in feats]
    t = cm2.prepare(meta).to(dev)
    pl, dl = cm2.latents_from_pooled(*[f.mean(1) for f in full], t)
    ref = cm2.loss_from_latents(pl, dl, t.G)
    ref.backward()
    ok = abs(float(loss) - float(ref)) <= 1e-5 * max(1.0, abs(float(ref)))
    gerr = 0.0
    for gl, f in zip(loc, full):
        want = f.grad[rank * B:(rank + 1) * B] * world
        gerr = max(gerr, float((gl.grad - want).abs().max() / (want.abs().max() + 1e-12)))
    ok = ok and gerr <= 1e-3 and float(ref) > 0
    res = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    print(f"rank {rank}/{world}: loss {float(loss):.6f} ref {float(ref):.6f} max grad rel err {gerr:.2e} "
          f"{'OK' if ok else 'MISMATCH'}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if float(res) == 1.0 else 1)


if __name__ == "__main__":
    main()
