#!/bin/bash
# N-rank evidence: sharded step == global batch, NCCL halo exchange == whole graph, bench at N ranks.  usage: call_o.sh N
n=$1; o=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29543 tests/multi/sharded_step_parity.py $o/r02_sharded_step_parity_${n}gpu.jsonl 2>&1 | grep -E "SHARDED|Error|error" | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29553 tests/multi/halo_parity.py $o/r02_halo_parity_${n}gpu.jsonl 2>&1 | grep -E "HALO|Error|error" | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > $o/r02_bench_${n}gpu_e.json 2> $o/r02_bench_${n}gpu_e.err; echo "bench rc=$?"
python - <<P
import json
d=json.load(open('gpurun_out/r02_bench_${n}gpu_e.json')); print(d['ms_per_step'], d['value'], d.get('parity_probe'), {k: (d[k].get('ms_per_step') if isinstance(d.get(k), dict) else None) for k in ('cylinder','c4')})
P
