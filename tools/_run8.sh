cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in "early:DL_EARLY_UPDATE=1" "noovl:DL_NO_OVERLAP=1"; do
tag=${v%%:*}; envs=${v#*:}
env $envs timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2w_bench_8gpu_$tag.json 2> gpurun_out/r2w_bench_8gpu_$tag.err; grep "^{" gpurun_out/r2w_bench_8gpu_$tag.json | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('8gpu $tag', d['value'], d['ms_per_step'])" || tail -5 gpurun_out/r2w_bench_8gpu_$tag.err
done
