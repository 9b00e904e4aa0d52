cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 100 python tools/mhla_gemm_probe.py
DL_GEMM_CTA2=0 timeout 100 python tools/mhla_gemm_probe.py
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 4 -c 2 -f -o gpurun_out/r2w_mhla python tools/mhla_gemm_probe.py --once > /dev/null 2>&1; ls -la gpurun_out/r2w_mhla.ncu-rep
