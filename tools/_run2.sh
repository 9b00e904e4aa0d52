cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for e in 0 1; do
DL_EARLY_UPDATE=$e timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$e bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r2w_bench_2gpu_early$e.json 2> gpurun_out/r2w_bench_2gpu_early$e.err; grep "^{" gpurun_out/r2w_bench_2gpu_early$e.json | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('2gpu early=$e', d['value'], d['ms_per_step'], d['loss'])" || tail -5 gpurun_out/r2w_bench_2gpu_early$e.err
done
