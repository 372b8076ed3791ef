#!/bin/bash
# final two-rank check of the committed build: parity scripts + the bench line the driver will produce
o=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tests/multi/sharded_step_parity.py $o/r02_sharded_step_parity_2gpu_final.jsonl 2>&1 | grep -E "SHARDED|Error|error" | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > $o/r02_bench_2gpu_final.json 2> $o/r02_bench_2gpu_final.err; echo "bench rc=$?"
python - <<P
import json
d = json.load(open('gpurun_out/r02_bench_2gpu_final.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['parity_probe']['grad_rel_all'], d['cylinder']['ms_per_step'], d['c4']['ms_per_step'], d['c4'].get('efficiency_vs_n1'))
P
