#!/bin/bash
o=gpurun_out
for cfg in 74,74 66,82; do
MMPDE_BRANCH_SMS=$cfg MMPDE_KINETO=$o/r02_kineto_w_$cfg.txt timeout 600 python bench.py --steps 40 --warmup 5 --no-extras --no-cpu-baseline > $o/r02_bench_w_$cfg.json 2> $o/r02_bench_w_$cfg.err; echo "uniform,moved = $cfg: rc=$?"; python -c "
import json; d=json.load(open('$o/r02_bench_w_$cfg.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['rollout']['ms_per_step'])"; sed -n 2,3p $o/r02_kineto_w_$cfg.txt; done
