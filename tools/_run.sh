set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_r2w.txt 2>&1; tail -4 gpurun_out/pytest_r2w.txt
timeout 300 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench_r2w.json 2> gpurun_out/bench_r2w.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r2w.json'));print('EARLY', d['value'], d['ms_per_step'], d['gpu_launches_per_step'])"
DL_NO_EARLY_UPDATE=1 timeout 300 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench_r2w_noearly.json 2> gpurun_out/bench_r2w.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r2w_noearly.json'));print('NOEARLY', d['value'], d['ms_per_step'], d['gpu_launches_per_step'])"
