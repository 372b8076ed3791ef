#!/bin/bash
# two ranks: branch priority off / on
o=gpurun_out
for v in 0 1; do
MMPDE_BRANCH_PRIORITY=$v MMPDE_KINETO=$o/r02_kineto_bench_2gpu_prio$v.txt timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras --no-cpu-baseline > $o/r02_bench_2gpu_prio$v.json 2> $o/r02_bench_2gpu_prio$v.err; echo "prio=$v rc=$?"
python - <<P
import json
d = json.load(open('gpurun_out/r02_bench_2gpu_prio$v.json')); print($v, d['ms_per_step'], d['e2e']['ms_per_step'])
P
grep -E "bn_stats|bn_bwd_reduce|^step" $o/r02_kineto_bench_2gpu_prio$v.txt
done
