#!/bin/bash
o=gpurun_out
for cfg in "0 0" "2 2" "0 2" "2 3"; do set -- $cfg
MMPDE_FULL_FWD_UNIFORM=$1 MMPDE_FULL_BWD_MOVED=$2 timeout 600 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline > $o/r02_bench_w$1$2.json 2> $o/r02_bench_w$1$2.err; echo "full fwd uniform < $1, full bwd moved < $2: rc=$?"; python -c "
import json; d=json.load(open('$o/r02_bench_w$1$2.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'])"; done
