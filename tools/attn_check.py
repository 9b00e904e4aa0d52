"""Bring-up / regression check of dl_attn_fwd / dl_attn_bwd against a plain fp32 PyTorch attention on
the same bf16 inputs, one subprocess per shape (a trapped kernel poisons its CUDA context only).

    python tools/attn_check.py            # all cases
    python tools/attn_check.py --case 3   # one case, in-process
    python tools/attn_check.py --time     # also time fwd / bwd (CUDA events)
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (name, B, H, S2, Lq, Lk, d, raw, stride mode)
CASES = [
    ("tiny d64 1 tile", 1, 1, 1, 128, 128, 64, False, "dense"),
    ("d64 Lk256", 2, 2, 1, 128, 256, 64, False, "dense"),
    ("paired PMMA (fused qkv views)", 3, 4, 2, 256, 256, 64, False, "qkv"),
    ("plain PMMA d128", 2, 4, 1, 256, 256, 128, False, "qkv"),
    ("PGCA model shape + raw", 2, 1, 1, 256, 512, 128, True, "kv"),
    ("PGCA long-seq shape (ragged) + raw", 2, 1, 1, 300, 290, 128, True, "kv"),
    ("ragged small", 2, 1, 1, 77, 33, 128, True, "dense"),
    ("d64 4 chunks", 1, 2, 2, 200, 400, 64, False, "dense"),
    # the step's real sizes (64 pairs) and the long-sequence config (256 pairs): timing cases
    ("FULL paired PMMA", 64, 4, 2, 256, 256, 64, False, "qkv"),
    ("FULL plain PMMA", 64, 4, 1, 256, 256, 128, False, "qkv"),
    ("FULL PGCA", 64, 1, 1, 256, 512, 128, True, "kv"),
    ("FULL PGCA long-seq B=256", 256, 1, 1, 1200, 290, 128, True, "kv"),
]


def run_case(idx: int, timing: bool) -> int:
    import torch

    from druglamp_b200 import kernels as K

    name, B, H, S2, Lq, Lk, d, want_raw, mode = CASES[idx]
    HD = H * d
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(100 + idx)

    def rnd(*shape, s=1.0):
        return (torch.randn(*shape, generator=g) * s).to(dev).to(torch.bfloat16)

    if mode == "qkv":          # column slices of a fused (S2, B, L, 3*HD) projection buffer
        assert Lq == Lk
        buf = rnd(S2, B, Lq, 3 * HD)
        q, k, v = buf[:, :, :, :HD], buf[0, :, :, HD:2 * HD], buf[0, :, :, 2 * HD:]
        dbuf = torch.zeros_like(buf)
        dq, dk, dv = dbuf[:, :, :, :HD], dbuf[0, :, :, HD:2 * HD], dbuf[0, :, :, 2 * HD:]
    elif mode == "kv":         # K and V are the column halves of one (B, Lk, 2*HD) buffer
        q = rnd(S2, B, Lq, HD)
        kv = rnd(B, Lk, 2 * HD)
        k, v = kv[:, :, :HD], kv[:, :, HD:]
        dq = torch.zeros_like(q)
        dkv = torch.zeros_like(kv)
        dk, dv = dkv[:, :, :HD], dkv[:, :, HD:]
    else:
        q, k, v = rnd(S2, B, Lq, HD), rnd(B, Lk, HD), rnd(B, Lk, HD)
        dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
    scale = float(d) ** -0.5
    d_o = rnd(B, Lq, S2 * HD)

    # fp32 reference on the same bf16 values
    qf = q.float().view(S2, B, Lq, H, d).permute(1, 3, 0, 2, 4).requires_grad_(True)   # (B,H,S2,Lq,d)
    kf = k.float().view(B, Lk, H, d).permute(0, 2, 1, 3).requires_grad_(True)         # (B,H,Lk,d)
    vf = v.float().view(B, Lk, H, d).permute(0, 2, 1, 3).requires_grad_(True)
    qf.retain_grad(); kf.retain_grad(); vf.retain_grad()
    s = torch.einsum("bhsqd,bhkd->bhsqk", qf, kf) * scale
    pr = torch.softmax(s, -1)
    of = torch.einsum("bhsqk,bhkd->bhsqd", pr, vf)                                    # (B,H,S2,Lq,d)
    o_ref = of.permute(0, 3, 2, 1, 4).reshape(B, Lq, S2 * HD)
    lse_ref = (torch.logsumexp(s, -1) * 1.4426950408889634).permute(2, 0, 1, 3)        # (S2,B,H,Lq)
    o_ref.backward(d_o.float())
    dq_ref = qf.grad.permute(2, 0, 3, 1, 4).reshape(S2, B, Lq, HD)
    dk_ref = kf.grad.permute(0, 2, 1, 3).reshape(B, Lk, HD)
    dv_ref = vf.grad.permute(0, 2, 1, 3).reshape(B, Lk, HD)

    o, lse, raw = K.attn_fwd(q, k, v, H, scale, want_raw=want_raw)
    torch.cuda.synchronize()

    def rel(a, b):
        return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-20))

    res = {"O": rel(o, o_ref), "lse": float((lse - lse_ref).abs().max())}
    if want_raw:
        res["raw"] = rel(raw, s[:, :, 0])
    K.attn_bwd(d_o, q, k, v, o, lse, H, scale, dq, dk, dv)
    torch.cuda.synchronize()
    res["dq"], res["dk"], res["dv"] = rel(dq, dq_ref), rel(dk, dk_ref), rel(dv, dv_ref)
    # accumulate mode: a second backward adds the same dq again
    K.attn_bwd(d_o, q, k, v, o, lse, H, scale, dq, dk, dv, dq_accumulate=True)
    torch.cuda.synchronize()
    res["dq_acc"] = rel(dq, 2 * dq_ref)
    bad = [k_ for k_, e in res.items() if not (e < (3e-2 if k_ != "lse" else 2e-2))]
    line = f"case {idx} [{name}] B={B} H={H} S2={S2} Lq={Lq} Lk={Lk} d={d}: " + \
           " ".join(f"{k_}={e:.2e}" for k_, e in res.items()) + ("  FAIL " + ",".join(bad) if bad else "  ok")
    if timing:
        def t(fn, n=20):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n * 1e3
        tf = t(lambda: K.attn_fwd(q, k, v, H, scale, want_raw=want_raw))
        tb = t(lambda: K.attn_bwd(d_o, q, k, v, o, lse, H, scale, dq, dk, dv))
        line += f"  fwd {tf:.1f} us bwd {tb:.1f} us"
    print(line, flush=True)
    return 1 if bad else 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", type=int, default=-1)
    ap.add_argument("--time", action="store_true")
    a = ap.parse_args()
    if a.case >= 0:
        return run_case(a.case, a.time)
    rc = 0
    for i in range(len(CASES)):
        cmd = [sys.executable, os.path.abspath(__file__), "--case", str(i)] + (["--time"] if a.time else [])
        try:
            r = subprocess.run(cmd, timeout=180, capture_output=True, text=True)
            out = (r.stdout + r.stderr).strip().splitlines()
            print("\n".join(out[-6:]) if r.returncode not in (0, 1) else r.stdout.strip(), flush=True)
            rc |= r.returncode != 0
        except subprocess.TimeoutExpired:
            print(f"case {i} [{CASES[i][0]}]: TIMEOUT", flush=True)
            rc = 1
    return rc


if __name__ == "__main__":
    sys.exit(main())
