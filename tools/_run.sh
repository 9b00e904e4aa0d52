cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py -m gpu -q -x -k "smallk or mhla or MHLA or Multi" 2>&1 | grep -v "^$" | tail -4
timeout 300 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/bench_A.json 2> gpurun_out/bench_A.err
python -c "
import json;d=json.load(open('gpurun_out/bench_A.json'));print('A', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['gemm_ms_per_step'], d['loss'])" || tail -5 gpurun_out/bench_A.err
