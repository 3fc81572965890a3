# final ncu captures of the round: the autoencoder layer kernel as shipped (split accumulators, narrow last round) and the
# bulk-copy-ring forward kernel
set -x
python tools/ncu_mlp.py > gpurun_out/plain_mlp.log 2>&1 && ncu --set full --clock-control none -k regex:mlp_gemm -s 1 -c 1 -o gpurun_out/r2_mlp_gemm_final python tools/ncu_mlp.py > gpurun_out/ncu_mlp.log 2>&1
tail -1 gpurun_out/ncu_mlp.log
python tools/ncu_shapes.py forward35 > gpurun_out/plain_fwd.log 2>&1 && ncu --set full --clock-control none -k regex:forward_tma -s 1 -c 1 -o gpurun_out/r2_forward_tma35 python tools/ncu_shapes.py forward35 > gpurun_out/ncu_fwd.log 2>&1
tail -1 gpurun_out/ncu_fwd.log
ls -la gpurun_out/*.ncu-rep
