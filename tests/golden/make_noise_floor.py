"""bf16 noise floor of the train-mode fixtures, measured on the UNMODIFIED reference.

    PYTHONPATH=/root/repo python tests/golden/make_noise_floor.py

Train-mode BatchNorm over a small batch amplifies rounding noise: PyTorch's own bf16 autocast run
of the reference moves the logits of these fixtures by O(1) of their range.  A single autocast run
is one draw from that noise; the product's bf16 path (different but equally valid rounding points,
tanh-form GELU on bf16 activations) is another.  This script draws an ensemble: the reference under
bf16 autocast with every fp32 parameter perturbed by a relative U(-2^-9, 2^-9) -- half a bf16 ulp,
the rounding any bf16 implementation applies to the weights -- and records the spread of the loss
and logits against the fp32 reference.  tests/test_model_parity_gpu.py bounds the product's bf16
deviation on these fixtures by the ensemble maximum (x1.5), not by a hand-picked constant.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402
from druglamp_b200.synth import make_batch  # noqa: E402
from tests.golden.make_golden import build, load_det  # noqa: E402

N_DRAWS = 12
CASES = [("DrugLAMP2C2P", 16, 7, "druglamp2c2p_train_b16"), ("DrugLAMPwoLLM", 12, 5, "druglampwollm_train_b12")]


def forward(m, b, B, autocast):
    from model.basic_model import binary_cross_entropy
    g = ref_shim.FakeGraph(b.graph.src, b.graph.dst, b.graph.num_nodes(), B, b.graph.ndata["h"].clone())
    with torch.no_grad():
        if autocast:
            with torch.autocast("cpu", dtype=torch.bfloat16):
                s = m(g, b.vp, b.xd, b.xp)[4].float()
        else:
            s = m(g, b.vp, b.xd, b.xp)[4].float()
    _, loss = binary_cross_entropy(s, b.y)
    return s, float(loss)


def main():
    out = {}
    for kind, B, seed, name in CASES:
        fx = np.load(os.path.join(HERE, name + ".npz"))
        assert int(fx["meta_B"]) == B and int(fx["meta_seed"]) == seed, "case table out of date"
        b = make_batch(B, seed=seed)
        m = build(kind)
        load_det(m)
        m.train(True)
        s32, l32 = forward(m, b, B, False)
        assert abs(l32 - float(fx["loss"])) < 1e-5 * abs(l32), (l32, float(fx["loss"]))
        base = {k: v.clone() for k, v in m.state_dict().items()}
        ldev, sdev = [], []
        for d in range(N_DRAWS):
            gen = torch.Generator().manual_seed(1000 + d)
            sd = {}
            for k, v in base.items():
                if v.dtype == torch.float32 and "running_" not in k:
                    sd[k] = v * (1.0 + (torch.rand(v.shape, generator=gen) * 2 - 1) * 2.0 ** -9)
                else:
                    sd[k] = v.clone()
            m.load_state_dict(sd, strict=True)
            m.train(True)
            s16, l16 = forward(m, b, B, True)
            ldev.append(abs(l16 - l32) / abs(l32))
            sdev.append(float((s16 - s32).abs().max() / s32.abs().max()))
            print(f"{name} draw {d}: loss dev {ldev[-1]:.4f} score dev {sdev[-1]:.4f}", flush=True)
        out[name + "/loss_dev"] = np.array(ldev)
        out[name + "/score_dev"] = np.array(sdev)
    np.savez(os.path.join(HERE, "bf16_noise_floor.npz"), **out)


if __name__ == "__main__":
    main()
