cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/nccl_cm_check.py > gpurun_out/r2w_nccl_cm_check_8gpu.txt 2>&1; grep -v "Warning\|warn\|OMP_NUM\|\*\*\*\*" gpurun_out/r2w_nccl_cm_check_8gpu.txt | tail -6
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r2w_bench_8gpu.json 2> gpurun_out/r2w_bench_8gpu.err; grep "^{" gpurun_out/r2w_bench_8gpu.json | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('8gpu', d['value'], d['ms_per_step'], d['e2e']['value'])" || tail -5 gpurun_out/r2w_bench_8gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --config 2c2p --steps 3 --warmup 2 > gpurun_out/r2w_bench_2c2p_8gpu.json 2>> gpurun_out/r2w_bench_8gpu.err; grep "^{" gpurun_out/r2w_bench_2c2p_8gpu.json | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('2c2p 8gpu', d['value'], d['ms_per_step'])" || tail -5 gpurun_out/r2w_bench_8gpu.err
