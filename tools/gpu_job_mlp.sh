set -x
timeout 600 python -m pytest tests/test_gpu_mlp.py -x -q 2>&1 | tail -25
timeout 600 python symmetry-ode-discovery_b200/sindy_b200/run.py --reference baseline/_ref tools/time_c3_closure.py 2>&1 | tail -60 > gpurun_out/c3_closure.txt
cat gpurun_out/c3_closure.txt
timeout 1500 python -m pytest tests/test_gpu_configs.py -x -q -k "config3" -s 2>&1 | tail -8
