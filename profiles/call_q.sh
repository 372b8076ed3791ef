#!/bin/bash
o=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_properties.py tests/test_gpu_fullsize.py tests/test_gpu_path.py -q -m gpu -x -k "knn or itp or interp or graph or create or c4 or C4 or c5 or cylinder or radius or dmm" > $o/r02_pytest_knn.log 2>&1; echo "tests rc=$?"; tail -3 $o/r02_pytest_knn.log
timeout 600 python bench.py --steps 20 --warmup 5 > $o/r02_bench_g.json 2> $o/r02_bench_g.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$o/r02_bench_g.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['rollout']['ms_per_step'], d['cylinder']['ms_per_step']); print([ (k['name'], round(k['avg_us'],1)) for k in d['kernels'] if 'knn' in k['name'] or 'dmm' in k['name']])"
