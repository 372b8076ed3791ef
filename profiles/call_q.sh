#!/bin/bash
o=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_guards.py -q -m gpu -x -k "wgrad or layer or solver or decoder or node" > $o/r02_pytest_wgrad.log 2>&1; echo "tests rc=$?"; tail -2 $o/r02_pytest_wgrad.log
timeout 300 python profiles/edge_bench.py 30 2>&1 | grep node_
timeout 600 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline > $o/r02_bench_h.json 2> $o/r02_bench_h.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$o/r02_bench_h.json')); print(d['ms_per_step'], d['e2e']['ms_per_step']); print([ (k['name'], round(k['avg_us'],1)) for k in d['kernels'] if 'wgrad' in k['name']])"
