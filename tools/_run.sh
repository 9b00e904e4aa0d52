cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x > gpurun_out/r2w_pytest_gpu.txt 2>&1; tail -2 gpurun_out/r2w_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; python -c "
import json;d=json.load(open('gpurun_out/r2w_bench.json'));print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['whole_step_frac'])" || tail -5 gpurun_out/r2w_bench.err
