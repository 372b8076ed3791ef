#!/bin/bash
# last check of the committed build: smoke + the bench line with default flags (what the driver runs)
o=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > $o/r02_bench_final.json 2> $o/r02_bench_final.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$o/r02_bench_final.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['rollout']['ms_per_step'], d['cylinder']['ms_per_step'], d['c4']['ms_per_step'], d['roofline']['frac'], d['cpu_baseline']['value'], d['gpu_launches']); print(d['config']['launch'])"
