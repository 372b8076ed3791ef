#!/bin/bash
o=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > $o/r02_pytest_gpu_f.log 2>&1; echo "suite rc=$?"; tail -2 $o/r02_pytest_gpu_f.log
for b in "" 0; do MMPDE_BRANCH_SMS=$b; if [ -z "$b" ]; then unset MMPDE_BRANCH_SMS; else export MMPDE_BRANCH_SMS; fi
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $o/r02_bench_i$b.json 2> $o/r02_bench_i$b.err; echo "branch_sms='$b' rc=$?"; python -c "
import json; d=json.load(open('$o/r02_bench_i$b.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['rollout']['ms_per_step'], d['cylinder']['ms_per_step'], d['roofline']['frac'])"; done
