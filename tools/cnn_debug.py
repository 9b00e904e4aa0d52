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
    out = m(tok.cuda(), fill.cuda())
    out.backward(dy.cuda())
    torch.cuda.synchronize()

    for dev in ("cuda", "cpu"):
        p = {k: v.clone().to(dev).requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
        x = torch.cat((F.embedding(tok.long().to(dev), p["embedding.weight"]), fill.to(dev).unsqueeze(-1)), -1).transpose(2, 1)
        for i in (1, 2, 3):
            x = F.relu(F.conv1d(x, p[f"conv{i}.weight"], p[f"conv{i}.bias"], padding="same"))
            x = F.batch_norm(x, None, None, p[f"bn{i}.weight"], p[f"bn{i}.bias"], True, 0.1, 1e-5)
        ref = x.reshape(B, L, 128)
        ref.backward(dy.to(dev))
        print(f"--- torch fp32 on {dev}: fwd err {float((out.float().to(dev) - ref).abs().max() / ref.abs().max()):.2e}")
        for k, v in m.named_parameters():
            a, r = v.grad.double().flatten().to(dev), p[k].grad.double().flatten()
            print(f"  {k:20s} max-err/max {float((a - r).abs().max() / r.abs().max()):.2e}  rel-L2 {float((a - r).norm() / r.norm()):.2e}")


if __name__ == "__main__":
    main()
