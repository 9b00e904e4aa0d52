"""CPU, world_size 2, gloo: the host-side logic of the two exchange steps.

* CrossModality with all-gathered negatives == one process on the rank-ordered concatenated batch
  (loss, and the gradient reaching each rank's local features), with the device arithmetic replaced
  by the CPU oracle (the dl_* kernels need a GPU; what is under test is the gather / dedup /
  gradient-routing logic of druglamp_b200/parallel.py).
* The flat-parameter gradient all-reduce == the mean of the per-rank gradients."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_cm_patch(cm, sd, margin):
    """Replace the kernel-backed pieces of CrossModality with the CPU oracle (same math)."""
    from oracle import restatement as R

    def m2e(name, x):
        w, b = sd[f"{name}.0.weight"], sd[f"{name}.0.bias"]
        y = F.batch_norm(x, None, None, w, b, True, 0.1, 1e-5)
        return F.linear(F.relu(y), sd[f"{name}.2.weight"], sd[f"{name}.2.bias"])

    def latents_from_pooled(prot, aug_prot, drug, aug_drug, targets):
        pe = torch.cat((m2e("prot2latent", prot[targets.p_idx]), m2e("aug_prot2latent", aug_prot[targets.p_idx])), -1)
        de = torch.cat((m2e("drug2latent", drug[targets.d_idx]), m2e("aug_drug2latent", aug_drug[targets.d_idx])), -1)
        return (F.normalize(F.linear(pe, sd["to_prot_latent.weight"]), dim=-1),
                F.normalize(F.linear(de, sd["to_drug_latent.weight"]), dim=-1))

    cm.latents_from_pooled = latents_from_pooled
    cm.loss_from_latents = lambda pl, dl, G: R.cm_triplet_dense(pl, dl, G.long(), margin)


def _make_inputs(B, seed):
    from druglamp_b200.synth import make_batch
    g = torch.Generator().manual_seed(seed)
    meta = make_batch(B, seed=seed, drugs_per_protein=3.0).meta
    feats = [torch.randn(B, 6, 128, generator=g) for _ in range(4)]
    return feats, meta


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from druglamp_b200.modules import CrossModality
        from druglamp_b200.parallel import global_cross_modality_loss
        torch.manual_seed(0)
        cm = CrossModality(hidden_size=128)
        sd = {k: v.detach().clone() for k, v in cm.state_dict().items()}
        _oracle_cm_patch(cm, sd, 0.5)
        B = 12
        feats, meta = _make_inputs(B * world, seed=21)
        loc = [f[rank * B:(rank + 1) * B].clone().requires_grad_(True) for f in feats]
        loss = global_cross_modality_loss(cm, *loc, meta[rank * B:(rank + 1) * B],
                                          pool_fn=lambda s: s.mean(1))
        loss.backward()
        # flat-parameter gradient all-reduce on a tiny CPU module
        from druglamp_b200.params import FlatParams
        torch.manual_seed(1)
        net = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.Linear(8, 2))
        flat = FlatParams(net)
        x = torch.randn(5, 8, generator=torch.Generator().manual_seed(100 + rank))
        flat.zero_grad()
        net(x).square().mean().backward()
        local = flat.grad.clone()
        dist.all_reduce(flat.grad)
        flat.grad /= world
        ret[rank] = dict(loss=float(loss), grads=[t.grad.clone() for t in loc], flat_local=local,
                         flat_mean=flat.grad.clone(), views_ok=all(p.grad.data_ptr() >= flat.grad.data_ptr() for p in net.parameters()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_cm_all_gather_and_flat_grad_allreduce():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    # single-process reference on the concatenated batch
    from druglamp_b200.modules import CrossModality
    torch.manual_seed(0)
    cm = CrossModality(hidden_size=128)
    sd = {k: v.detach().clone() for k, v in cm.state_dict().items()}
    _oracle_cm_patch(cm, sd, 0.5)
    B = 12
    feats, meta = _make_inputs(B * world, seed=21)
    full = [f.clone().requires_grad_(True) for f in feats]
    t = cm.prepare(meta)
    pl, dl = cm.latents_from_pooled(*[f.mean(1) for f in full], t)
    ref = cm.loss_from_latents(pl, dl, t.G)
    ref.backward()
    assert ref.item() > 0
    for r in range(world):
        assert abs(ret[r]["loss"] - ref.item()) < 1e-6, (ret[r]["loss"], ref.item())
        for g_loc, f in zip(ret[r]["grads"], full):
            # averaged-DDP convention: local slice gradient is scaled by world_size
            assert torch.allclose(g_loc, f.grad[r * B:(r + 1) * B] * world, atol=1e-7, rtol=1e-5)
        assert ret[r]["views_ok"]
    mean = (ret[0]["flat_local"] + ret[1]["flat_local"]) / 2
    assert torch.allclose(ret[0]["flat_mean"], mean) and torch.allclose(ret[1]["flat_mean"], mean)
