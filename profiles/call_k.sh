#!/bin/bash
o=gpurun_out
timeout 900 python -m pytest tests/test_gpu_guards.py tests/test_gpu_kernels.py -q -m gpu -x 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_path.py tests/test_gpu_properties.py -q -m gpu -x 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $o/r02_bench_k.json 2> $o/r02_bench_k.err; echo "bench rc=$?"; cut -c1-260 $o/r02_bench_k.json; tail -3 $o/r02_bench_k.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench_k.json'))
for r in d['kernels']: print(r['name'], r['launches_per_step'], round(r['avg_us'],1), round(r['min_us'],1), round(r['us_per_step']))
P
