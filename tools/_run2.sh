cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2v_bench_2gpu.json 2> gpurun_out/r2v_bench_2gpu.err; python -c "
import json;d=json.load(open('gpurun_out/r2v_bench_2gpu.json'));print('2gpu', d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus'])" || tail -5 gpurun_out/r2v_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/r2v_bench_ref_2gpu.json 2>> gpurun_out/r2v_bench_2gpu.err; cut -c1-160 gpurun_out/r2v_bench_ref_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --config 2c2p --steps 3 --warmup 2 > gpurun_out/r2v_bench_2c2p_2gpu.json 2>> gpurun_out/r2v_bench_2gpu.err; python -c "
import json;d=json.load(open('gpurun_out/r2v_bench_2c2p_2gpu.json'));print('2c2p 2gpu', d['value'], d['ms_per_step'])" || tail -5 gpurun_out/r2v_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 tools/ddp_check.py 2>&1 | tail -3
