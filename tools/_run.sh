cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ffn_fused_gpu.py tests/test_gemm_gpu.py -m gpu -q -x 2>&1 | grep -v "^$" | tail -3
timeout 120 python tools/ffn_probe.py
for v in "A:DL_GEMM_CTA2=1" "B:DL_GEMM_CTA2=0"; do
  tag=${v%%:*}; envs=${v#*:}
  env $envs DL_BENCH_DUMP=gpurun_out/gemm_profile_$tag.txt timeout 300 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python -c "
import json;d=json.load(open('gpurun_out/bench_$tag.json'));print('$tag $envs', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['gemm_ms_per_step'], d['loss'])" || tail -5 gpurun_out/bench_$tag.err
done
