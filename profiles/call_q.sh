#!/bin/bash
o=gpurun_out
timeout 600 python -m pytest tests/test_gpu_path.py -q -m gpu -x -k "dmm or displacement or golden or g6 or mover" -s > $o/r02_pytest_dmm.log 2>&1; echo "tests rc=$?"; tail -4 $o/r02_pytest_dmm.log
timeout 600 python bench.py --steps 20 --warmup 5 > $o/r02_bench_f.json 2> $o/r02_bench_f.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$o/r02_bench_f.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['rollout']['ms_per_step'], d['cylinder']['ms_per_step'])"
