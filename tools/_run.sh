set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_r2v.txt 2>&1; tail -4 gpurun_out/pytest_r2v.txt
timeout 300 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench_r2v.json 2> gpurun_out/bench_r2v.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r2v.json'));print('2WG', d['value'], d['ms_per_step'], d['gpu_launches_per_step'])"
DL_WGRAD_STREAMS=3 timeout 300 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench_r2v3.json 2> gpurun_out/bench_r2v.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r2v3.json'));print('3WG', d['value'], d['ms_per_step'], d['gpu_launches_per_step'])"
DL_WGRAD_STREAMS=1 timeout 300 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench_r2v1.json 2> gpurun_out/bench_r2v.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r2v1.json'));print('1WG', d['value'], d['ms_per_step'], d['gpu_launches_per_step'])"
