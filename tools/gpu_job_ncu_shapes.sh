python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_gpu_suite.log; tail -6 gpurun_out/r2_gpu_suite.log
python symmetry-ode-discovery_b200/sindy_b200/run.py --reference baseline/_ref tools/time_c3_closure.py 20000 2>&1 | tail -40 > gpurun_out/r2_c3_closure.log; cat gpurun_out/r2_c3_closure.log
for w in fused33 fused23 moments35 forward35; do
  python tools/ncu_shapes.py $w > gpurun_out/plain_$w.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"fused_step|moments_kernel|forward_spec" -s 1 -c 1 -o gpurun_out/r2_$w python tools/ncu_shapes.py $w > gpurun_out/ncu_$w.log 2>&1; tail -1 gpurun_out/ncu_$w.log
done
