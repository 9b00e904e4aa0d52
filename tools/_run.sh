set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ffn_fused_gpu.py -m gpu -q -x > gpurun_out/pytest_ffn.txt 2>&1; tail -5 gpurun_out/pytest_ffn.txt
timeout 120 python tools/ffn_probe.py > gpurun_out/ffn_probe.txt 2>&1; cat gpurun_out/ffn_probe.txt
