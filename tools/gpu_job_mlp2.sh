set -x
timeout 300 python tools/time_mlp.py 40000 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_mlp.py -x -q 2>&1 | tail -3
python tools/ncu_mlp.py > gpurun_out/plain_mlp.log 2>&1 && ncu --set full --clock-control none -k regex:mlp_gemm -s 1 -c 1 -o gpurun_out/r2_mlp_gemm python tools/ncu_mlp.py > gpurun_out/ncu_mlp.log 2>&1
tail -2 gpurun_out/ncu_mlp.log; ls -la gpurun_out/r2_mlp_gemm.ncu-rep
