set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2m.txt 2>&1; tail -3 gpurun_out/pytest_r2m.txt
timeout 300 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench_r2m.json 2> gpurun_out/bench_r2m.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r2m.json'));print('STREAMS', d['value'], d['ms_per_step'], d['e2e']['value'])"
DL_NO_BRANCH_STREAMS=1 timeout 300 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench_r2m_nostreams.json 2>> gpurun_out/bench_r2m.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r2m_nostreams.json'));print('NOSTREAMS', d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 300 python tools/graph_profile.py > gpurun_out/graph_profile_r2m.md 2>&1; head -3 gpurun_out/graph_profile_r2m.md
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/gemm_top_r2m python tools/gemm_ncu_top.py --top 10 > gpurun_out/gemm_top_r2m.txt 2>&1; tail -45 gpurun_out/gemm_top_r2m.txt
