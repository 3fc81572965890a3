set -x
export SB_MLP_PAIR=1
python tools/ncu_mlp.py > gpurun_out/plain_mlp_pair.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mlp_gemm -s 1 -c 1 -o gpurun_out/r2_mlp_gemm_pair python tools/ncu_mlp.py > gpurun_out/ncu_mlp_pair.log 2>&1
tail -2 gpurun_out/ncu_mlp_pair.log; ls -la gpurun_out/r2_mlp_gemm_pair.ncu-rep
