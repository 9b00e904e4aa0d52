set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2n.txt 2>&1; tail -3 gpurun_out/pytest_r2n.txt
timeout 300 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench_r2n.json 2> gpurun_out/bench_r2n.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r2n.json'));print('WGRAD-STREAM', d['value'], d['ms_per_step'], d['e2e']['value'])"
DL_NO_WGRAD_STREAM=1 timeout 300 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench_r2n_nowg.json 2>> gpurun_out/bench_r2n.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r2n_nowg.json'));print('NO-WGRAD-STREAM', d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 600 python tools/gemm_bench.py --splitk > gpurun_out/splitk_r2n.txt 2>&1; tail -60 gpurun_out/splitk_r2n.txt
