"""Bring-up probe for dl_gemm: runs each (dtype, layout, tile) group in its own process so a
trap in one configuration does not hide the others.  Writes gpurun_out/gemm_probe.txt."""
import itertools
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(dtype_s, ta, tb):
    import torch
    from tests.test_gemm_gpu import run_case
    dtype = {"bf16": torch.bfloat16, "f32": torch.float32}[dtype_s]
    out = []
    dead = False
    for tile_n in (128, 64, 256):
        for (M, N, K) in [(128, tile_n, 64), (128, tile_n, 256), (256, 256, 256), (200, 136, 328), (128, 8, 1024)]:
            try:
                err, scale, _ = run_case(M, N, K, dtype, ta, tb, tile_n=tile_n, out_dtype=torch.float32)
                out.append((tile_n, M, N, K, err, scale))
            except AssertionError as e:
                out.append((tile_n, M, N, K, repr(e)[:200], None))
            except Exception as e:  # noqa: BLE001
                out.append((tile_n, M, N, K, repr(e)[:300], None))
                dead = True
                break
        if dead:
            break
    print("RESULT " + json.dumps(out))


def main():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    lines = []
    for dtype_s, (ta, tb) in itertools.product(["bf16", "f32"], itertools.product([0, 1], [0, 1])):
        tile_n = "*"
        cmd = [sys.executable, __file__, "child", dtype_s, str(ta), str(tb)]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=180, cwd=ROOT)
            res = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            tail = (r.stdout + r.stderr)[-400:].replace("\n", " | ") if not res else ""
            line = f"{dtype_s} ta={ta} tb={tb} bn={tile_n}: {res[0][7:] if res else 'NO RESULT rc=%d %s' % (r.returncode, tail)}"
        except subprocess.TimeoutExpired:
            line = f"{dtype_s} ta={ta} tb={tb} bn={tile_n}: TIMEOUT"
        print(line, flush=True)
        lines.append(line)
    with open(os.path.join(ROOT, "gpurun_out", "gemm_probe.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(sys.argv[2], bool(int(sys.argv[3])), bool(int(sys.argv[4])))
    else:
        main()
