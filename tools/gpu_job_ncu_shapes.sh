# ncu captures of the config-shaped kernels (one launch each, full set, no source import to keep the reports small)
for w in fused33 fused23 moments35 forward35; do
  python tools/ncu_shapes.py $w > gpurun_out/plain_$w.log 2>&1 && ncu --set full --clock-control none -k regex:"fused_step|moments_kernel|forward_spec" -s 1 -c 1 -o gpurun_out/r2_$w python tools/ncu_shapes.py $w > gpurun_out/ncu_$w.log 2>&1; tail -1 gpurun_out/ncu_$w.log; ls -la gpurun_out/r2_$w.ncu-rep
done
python -m pytest tests/test_gpu_configs.py -m gpu -q -k "sweep" 2>&1 | tail -5
