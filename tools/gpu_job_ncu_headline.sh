# round-2 ncu evidence for the headline step: launch list of the bench command and one full capture of the fused kernel
set -x
B="python bench.py --steps 5 --warmup 3 --skip-extras --e2e-steps 1 --cpu-samples 50000"
timeout 300 $B > gpurun_out/r02_plain_bench.json 2> gpurun_out/r02_plain_bench.err || { tail -5 gpurun_out/r02_plain_bench.err; exit 1; }
tail -c 600 gpurun_out/r02_plain_bench.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_n1.csv $B > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_step_kernel -s 4 -c 1 -o gpurun_out/r02_fused35 $B > gpurun_out/ncu_fused35.log 2>&1
tail -2 gpurun_out/ncu_fused35.log
ls -la gpurun_out/r02_fused35.ncu-rep gpurun_out/r02_launches_bench_n1.csv
