cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_modules_gpu.py tests/test_kernels_gpu.py tests/test_model_parity_gpu.py -m gpu -q -x 2>&1 | grep -v "^$" | tail -3
for v in "A:DL_GEMM_CTA2=1"; do
  tag=${v%%:*}; envs=${v#*:}
  env $envs DL_BENCH_DUMP=gpurun_out/gemm_profile_$tag.txt timeout 300 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python -c "
import json;d=json.load(open('gpurun_out/bench_$tag.json'));print('$tag $envs', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['gemm_ms_per_step'], d['loss'])" || tail -5 gpurun_out/bench_$tag.err
done
timeout 200 python bench.py --config pgca --steps 50 2>/dev/null | cut -c1-200
