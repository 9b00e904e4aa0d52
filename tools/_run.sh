set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_r2t.txt 2>&1; tail -15 gpurun_out/pytest_r2t.txt
