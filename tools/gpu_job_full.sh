set -x
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02_gpu_suite.txt
cat gpurun_out/r02_gpu_suite.txt
timeout 900 python -m pytest tests/test_gpu_configs.py -q -s -k "config3" 2>&1 | grep -E "C3:|passed|failed" | tail -3
timeout 600 python symmetry-ode-discovery_b200/sindy_b200/run.py --reference baseline/_ref tools/time_c3_closure.py 2>&1 | grep -v "Warning\|warn\|Consider\|return Variable" > gpurun_out/c3_closure.txt
grep "===" gpurun_out/c3_closure.txt
timeout 900 python bench.py > gpurun_out/r02_bench_n1_new.json 2> gpurun_out/r02_bench_n1_new.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1_new.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'])
print(json.dumps(d['extra'].get('autoencoder_mlp_config3_40000rows'), indent=1))
PY
tail -3 gpurun_out/r02_bench_n1_new.err
