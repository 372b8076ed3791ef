#!/bin/bash
o=gpurun_out
for v in 0 1 0 1; do MMPDE_BRANCH_PRIORITY=$v timeout 600 python bench.py --steps 40 --warmup 5 > $o/r02_bench_prio$v.json 2> $o/r02_bench_prio$v.err; echo "prio=$v rc=$?"; python -c "
import json; d=json.load(open('$o/r02_bench_prio$v.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['rollout']['ms_per_step'], d['cylinder']['ms_per_step'])"; done
