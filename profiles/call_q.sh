#!/bin/bash
o=gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > $o/r02_bench_g.json 2> $o/r02_bench_g.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$o/r02_bench_g.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['rollout']['ms_per_step'], d['cylinder']['ms_per_step']); print([ (k['name'], round(k['avg_us'],1)) for k in d['kernels'] if 'knn' in k['name'] or 'dmm' in k['name']])"
