cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# one eager step: launch list + DRAM bytes + tensor-pipe activity per kernel
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/r2v_step_metrics.csv python bench.py --ncu-step > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log; wc -l gpurun_out/r2v_step_metrics.csv
# --set full of the kernels the step's time is in
DL_GEMM_CTA2=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/r2v_gemm_pair python tools/gemm_bench.py --one 16384,512,2048,1,1,1,0,0 > /dev/null 2>&1
DL_GEMM_CTA2=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/r2v_gemm_single python tools/gemm_bench.py --one 16384,512,2048,1,1,1,0,0 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ffn_chain -s 4 -c 2 -f -o gpurun_out/r2v_ffn python tools/ffn_probe.py --once > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:layernorm_bwd -s 1 -c 2 -f -o gpurun_out/r2v_lnbwd python tools/ln_bench.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
