"""Debug aid: the product's ProteinCNN (fp32 mode) against the same stack in plain torch fp32 on the
GPU, same weights / tokens / upstream gradient: per-parameter gradient error with full tensors."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def main():
    import druglamp_b200 as D
    from druglamp_b200.modules import ProteinCNN
    from oracle import restatement as R
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    D.set_compute_dtype(torch.float32)
    B, L = 12, 2304
    m = ProteinCNN(128, [128] * 3, [3, 6, 9], True).cuda()
    shapes = {"protein_extractor." + k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = {k[len("protein_extractor."):]: v for k, v in R.deterministic_state(shapes).items()}
    m.load_state_dict(sd)
    m.train()
    g = torch.Generator().manual_seed(0)
    tok = torch.randint(0, 26, (B, L), generator=g).double()
    tok[:, 1500:] = 0
    fill = (tok == 0).float()
    dy = torch.randn(B, L, 128, generator=g) * 1e-4
    # layer by layer, with hooks on the activations' gradients
    from druglamp_b200 import functions as Fn
    grads = {}
    emb = m.embedding
    x = Fn.EmbedFillFn.apply(tok.cuda(), fill.cuda(), emb.weight, emb.padding_idx)
    acts = {}
    for i in (1, 2, 3):
        conv, bn = getattr(m, f"conv{i}"), getattr(m, f"bn{i}")
        x = Fn.Conv1dSameFn.apply(x, conv.weight, conv.bias, True, False)
        acts[f"relu{i}"] = x.detach()
        x.register_hook(lambda g_, i=i: grads.__setitem__(f"d_relu{i}", g_.detach().clone()))
        x = Fn.batch_norm(x, bn, relu_input=True)
        x.register_hook(lambda g_, i=i: grads.__setitem__(f"d_bn{i}", g_.detach().clone()))
    y = Fn.TransposeFn.apply(x)
    out = y.view(y.size(0), y.size(2), -1)
    out.backward(dy.cuda())
    torch.cuda.synchronize()

    for dev in ("cuda", "cpu"):
        p = {k: v.clone().to(dev).requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
        x = torch.cat((F.embedding(tok.long().to(dev), p["embedding.weight"]), fill.to(dev).unsqueeze(-1)), -1).transpose(2, 1)
        racts = {}
        for i in (1, 2, 3):
            x = F.relu(F.conv1d(x, p[f"conv{i}.weight"], p[f"conv{i}.bias"], padding="same"))
            x.retain_grad(); racts[f"relu{i}"] = x
            x = F.batch_norm(x, None, None, p[f"bn{i}.weight"], p[f"bn{i}.bias"], True, 0.1, 1e-5)
            x.retain_grad(); racts[f"bn{i}"] = x
        ref = x.reshape(B, L, 128)
        ref.backward(dy.to(dev))
        for i in (3, 2, 1):
            # product activations are channels-last (B, L, C); the torch ones (B, C, L)
            rb = racts[f"bn{i}"].grad.transpose(1, 2)
            rr = (racts[f"relu{i}"].grad * (racts[f"relu{i}"] > 0)).transpose(1, 2)
            pb, pr = grads[f"d_bn{i}"].float().to(dev), grads[f"d_relu{i}"].float().to(dev)
            print(f"  layer {i}: d(bn out) rel-L2 {float((pb - rb).norm() / rb.norm()):.2e}   "
                  f"masked d(relu out) rel-L2 {float((pr - rr).norm() / rr.norm()):.2e}   "
                  f"act err {float((acts[f'relu{i}'].float().to(dev) - racts[f'relu{i}'].detach().transpose(1, 2)).abs().max()):.2e}")
        print(f"--- torch fp32 on {dev}: fwd err {float((out.float().to(dev) - ref).abs().max() / ref.abs().max()):.2e}")
        for k, v in m.named_parameters():
            a, r = v.grad.double().flatten().to(dev), p[k].grad.double().flatten()
            print(f"  {k:20s} max-err/max {float((a - r).abs().max() / r.abs().max()):.2e}  rel-L2 {float((a - r).norm() / r.norm()):.2e}")


if __name__ == "__main__":
    main()
