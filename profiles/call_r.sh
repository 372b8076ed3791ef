#!/bin/bash
# two ranks: BatchNorm backward exchange split around the deferred weight-gradient launch, off / on
o=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tests/multi/sharded_step_parity.py $o/r02_sharded_step_parity_split.jsonl 2>&1 | grep -E "SHARDED|Error|error" | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 tests/multi/halo_parity.py $o/r02_halo_parity_split.jsonl 2>&1 | grep -E "HALO|Error|error" | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -2
for v in 0 1; do
MMPDE_SPLIT_BN_EXCHANGE=$v MMPDE_KINETO=$o/r02_kineto_bench_2gpu_split$v.txt timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras --no-cpu-baseline > $o/r02_bench_2gpu_split$v.json 2> $o/r02_bench_2gpu_split$v.err; echo "split=$v rc=$?"
python - <<P
import json
d = json.load(open('gpurun_out/r02_bench_2gpu_split$v.json')); print($v, d['ms_per_step'], d['e2e']['ms_per_step'])
P
grep -E "bn_stats|bn_bwd_reduce|exchange_wait|^step" $o/r02_kineto_bench_2gpu_split$v.txt
done
