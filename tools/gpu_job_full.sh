set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02_gpu_suite.txt
cat gpurun_out/r02_gpu_suite.txt
timeout 600 python bench.py > gpurun_out/r02_bench_n1_new.json 2> gpurun_out/r02_bench_n1_new.err
tail -c 3000 gpurun_out/r02_bench_n1_new.json
tail -5 gpurun_out/r02_bench_n1_new.err
