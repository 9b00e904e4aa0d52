set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_model_parity_gpu.py tests/test_train_step_gpu.py -m gpu -q -x > gpurun_out/pytest_r2x.txt 2>&1; tail -4 gpurun_out/pytest_r2x.txt
DL_BENCH_DUMP=gpurun_out/gemm_profile_r2x.txt timeout 300 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench_r2x.json 2> gpurun_out/bench_r2x.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r2x.json'));print('F32X2', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['gemm_ms_per_step'])"
head -8 gpurun_out/gemm_profile_r2x.txt
