"""Multi-GPU check (torchrun, NCCL): the overlapped two-range gradient all-reduce inside the captured
step graph (train.TrainStep.overlap) against the plain single all-reduce after the backward.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/ddp_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    import druglamp_b200 as D
    from druglamp_b200.models import DrugLAMP
    from druglamp_b200.synth import make_batch
    from druglamp_b200.train import StaticBatch, TrainStep
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    D.set_compute_dtype(torch.bfloat16)
    res = {}
    for overlap in (True, False):
        os.environ["DL_NO_OVERLAP"] = "0" if overlap else "1"
        torch.manual_seed(1)
        m = DrugLAMP(384, 640).to(dev)
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        m.train()
        m.flatten_parameters()
        ts = TrainStep(m, world_size=world)
        assert ts.overlap == overlap
        sb = StaticBatch(make_batch(16, seed=50 + rank), dev)
        ts._fwd_bwd(sb)
        ts._reduce()
        torch.cuda.synchronize()
        eager = ts.flat.grad.clone()
        ts.capture(sb)
        for _ in range(3):
            loss = ts.replay(sb)
        torch.cuda.synchronize()
        res[overlap] = (eager, float(loss), ts.flat.flat.clone())
    scale = float(res[False][0].abs().max())
    d_grad = float((res[True][0] - res[False][0]).abs().max()) / scale
    d_par = float((res[True][2] - res[False][2]).abs().max())
    # all ranks must hold identical parameters after the replays
    ref = res[True][2].clone()
    dist.broadcast(ref, 0)
    same = float((ref - res[True][2]).abs().max())
    if rank == 0:
        print(f"world {world}: eager grad (overlapped vs plain all-reduce) rel diff {d_grad:.2e}; "
              f"params after 3 replays: overlap vs plain {d_par:.2e}, rank 0 vs rank {world - 1}... {same:.2e}; "
              f"losses {res[True][1]:.6f} / {res[False][1]:.6f}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
