"""Debug aid: per-parameter gradient error of the product (fp32 / bf16) against the CPU oracle on one
seeded batch, with COMPLETE tensors (max-abs error relative to the tensor's max, and cosine).

    python tools/grad_debug.py [kind] [B] [seed] [f32|bf16] [prefix]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "DrugLAMPwoLLM"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    mode = sys.argv[4] if len(sys.argv) > 4 else "f32"
    prefix = sys.argv[5] if len(sys.argv) > 5 else ""
    import druglamp_b200 as D
    from druglamp_b200 import models
    from druglamp_b200.modules import binary_cross_entropy
    from druglamp_b200.synth import make_batch
    from oracle import restatement as R

    D.set_compute_dtype(torch.float32 if mode == "f32" else torch.bfloat16)
    m = getattr(models, kind)(384, 640).cuda()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(R.deterministic_state(shapes), strict=True)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    m.train(True)
    b = make_batch(B, seed=seed)
    bc = b.to("cuda")
    out = m(*bc.model_inputs())
    _, loss = binary_cross_entropy(out[4], bc.y)
    loss.backward()
    torch.cuda.synchronize()

    torch.set_num_threads(os.cpu_count())
    with open(os.path.join(ROOT, "tests", "golden", "state_shapes.json")) as f:
        all_shapes = {k: tuple(v) for k, v in json.load(f).items()}
    sd = R.deterministic_state(all_shapes)
    leaves = {k: v.requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running_" not in k}
    o = R.druglamp_forward(sd, kind, b.graph.src, b.graph.dst, b.graph.ndata["h"], B, b.vp, b.xd, b.xp, True)
    _, rl = R.binary_cross_entropy(o["score"], b.y)
    rl.backward()
    print(f"{kind} B={B} {mode}: loss product {float(loss):.6f} oracle {float(rl):.6f}; "
          f"score err {float((out[4].float().cpu() - o['score']).abs().max() / o['score'].abs().max()):.2e}")
    rows = []
    for k, p in m.named_parameters():
        if p.grad is None or k not in leaves or leaves[k].grad is None or not k.startswith(prefix):
            continue
        g, r = p.grad.float().cpu().double().flatten(), leaves[k].grad.double().flatten()
        e = float((g - r).abs().max() / (r.abs().max() + 1e-30))
        c = float((g @ r) / (g.norm() * r.norm() + 1e-300))
        rows.append((e, k, c, float(r.abs().max()), float(r.abs().sum()), float(r.sum())))
    rows.sort(reverse=True)
    for e, k, c, mx, asum, s in rows[:25]:
        print(f"  {k:62s} err {e:.2e} cos {c:.6f} max {mx:.2e} abs-sum {asum:.2e} sum {s:.2e}")


if __name__ == "__main__":
    main()
