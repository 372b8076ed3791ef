#!/bin/bash
o=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > $o/r02_pytest_m.log 2>&1; echo "suite rc=$?"; tail -4 $o/r02_pytest_m.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tests/multi/sharded_step_parity.py $o/r02_sharded_step_parity_b.jsonl 2>&1 | grep -E "SHARDED|Error" | cut -c1-900
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 tests/multi/halo_parity.py $o/r02_halo_parity_b.jsonl 2>&1 | grep -E "HALO|Error" | cut -c1-1200
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $o/r02_bench_2gpu_m.json 2> $o/r02_bench_2gpu_m.err; echo "bench2 rc=$?"; cut -c1-200 $o/r02_bench_2gpu_m.json
python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench_2gpu_m.json')); print(d['ms_per_step'], d.get('parity_probe'))
P
