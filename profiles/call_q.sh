#!/bin/bash
# last check of the committed build
o=gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > $o/r02_pytest_gpu_g.log 2>&1; echo "suite rc=$?"; tail -1 $o/r02_pytest_gpu_g.log
timeout 900 python bench.py > $o/r02_bench_final2.json 2> $o/r02_bench_final2.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$o/r02_bench_final2.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['rollout']['ms_per_step'], d['cylinder']['ms_per_step'], d['c4']['ms_per_step'], d['roofline']['frac'])"
