set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "small_linear or head" > gpurun_out/pytest_head.txt 2>&1; tail -30 gpurun_out/pytest_head.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2r.txt 2>&1; tail -5 gpurun_out/pytest_r2r.txt
timeout 300 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench_r2r.json 2> gpurun_out/bench_r2r.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r2r.json'));print('HEAD', d['value'], d['ms_per_step'], d['gpu_launches_per_step'])"
DL_NO_HEAD_KERNELS=1 timeout 300 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench_r2r_nohead.json 2>> gpurun_out/bench_r2r.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r2r_nohead.json'));print('NOHEAD', d['value'], d['ms_per_step'], d['gpu_launches_per_step'])"
timeout 300 python tools/graph_profile.py gpurun_out/trace_r2r.json > gpurun_out/graph_profile_r2r.md 2>&1; head -3 gpurun_out/graph_profile_r2r.md
