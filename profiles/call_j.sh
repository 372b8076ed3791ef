#!/bin/bash
# 2-GPU call: guard tests, the real-rank parity tests, the 2-GPU bench
o=gpurun_out
timeout 600 python -m pytest tests/test_gpu_guards.py -q -m gpu 2>&1 | tail -8
timeout 1500 python -m pytest tests/test_gpu_multi.py -q -m gpu > $o/r02_pytest_multi_2gpu.log 2>&1; echo "multi rc=$?"; tail -5 $o/r02_pytest_multi_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $o/r02_bench_2gpu_j.json 2> $o/r02_bench_2gpu_j.err; echo "bench2 rc=$?"; cut -c1-220 $o/r02_bench_2gpu_j.json; tail -2 $o/r02_bench_2gpu_j.err
