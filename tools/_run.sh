set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --config 2c2p --steps 4 --warmup 2 > gpurun_out/r2_bench_2c2p_1gpu.json 2> gpurun_out/r2_2c2p.err; tail -c 600 gpurun_out/r2_bench_2c2p_1gpu.json
timeout 300 python bench.py --config pgca > gpurun_out/r2_bench_pgca_1gpu.json 2>> gpurun_out/r2_2c2p.err; tail -c 400 gpurun_out/r2_bench_pgca_1gpu.json
timeout 300 python bench.py --config infer > gpurun_out/r2_bench_infer_1gpu.json 2>> gpurun_out/r2_2c2p.err; tail -c 400 gpurun_out/r2_bench_infer_1gpu.json
DL_BENCH_DUMP=gpurun_out/gemm_profile_r2q.txt timeout 600 python bench.py > gpurun_out/bench_r2q.json 2> gpurun_out/bench_r2q.err; tail -c 1500 gpurun_out/bench_r2q.json
timeout 300 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/bench_ref_r2q.json 2>> gpurun_out/bench_r2q.err; cat gpurun_out/bench_ref_r2q.json | cut -c1-300
