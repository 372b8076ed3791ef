#!/bin/bash
# two ranks: persistent kernels at full width (branches take turns) vs half width (branches side by side)
o=gpurun_out
for b in 74 148; do
MMPDE_SM_BUDGET=$b MMPDE_KINETO=$o/r02_kineto_bench_2gpu_smb$b.txt timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras --no-cpu-baseline > $o/r02_bench_2gpu_smb$b.json 2> $o/r02_bench_2gpu_smb$b.err; echo "budget=$b rc=$?"
python - <<P
import json
d = json.load(open('gpurun_out/r02_bench_2gpu_smb$b.json')); print($b, d['ms_per_step'], d['e2e']['ms_per_step'])
P
grep -E "bn_stats|bn_bwd_reduce|exchange_wait|^step" $o/r02_kineto_bench_2gpu_smb$b.txt
done
