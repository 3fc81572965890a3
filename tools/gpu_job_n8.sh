# 8-GPU validation: exchange test at 8 ranks, then the bench at N = 8 (prints config.sharded_vs_unsharded)
set -x
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/peer_multi_check.py 2>&1 | tail -25 > gpurun_out/r02_peer_multi_check_n$N.txt
tail -12 gpurun_out/r02_peer_multi_check_n$N.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -c 1500 gpurun_out/r02_bench_n$N.json; tail -3 gpurun_out/r02_bench_n$N.err
