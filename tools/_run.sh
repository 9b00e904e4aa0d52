set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2p.txt 2>&1; tail -3 gpurun_out/pytest_r2p.txt
timeout 300 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench_r2p.json 2> gpurun_out/bench_r2p.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r2p.json'));print('PMMA-STREAMS+LN', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['gemm_ms_per_step'])"
timeout 300 python tools/graph_profile.py gpurun_out/trace_r2p.json > gpurun_out/graph_profile_r2p.md 2>&1; head -3 gpurun_out/graph_profile_r2p.md; grep layernorm_bwd gpurun_out/graph_profile_r2p.md
