#!/bin/bash
o=gpurun_out
timeout 600 python -m pytest "tests/test_gpu_path.py::test_training_and_rollout_loops_golden_fixture" tests/test_gpu_kernels.py -q -m gpu -k "golden_fixture or interpolation or res_cut" 2>&1 | tail -6
python profiles/debug/itp_ab.py 12 2>&1 | grep -E "round trip|moved.x|flow" 
timeout 1500 python -m pytest tests -q -m gpu > $o/r02_pytest_i.log 2>&1; echo "suite rc=$?"; tail -6 $o/r02_pytest_i.log
python bench.py --steps 10 --warmup 3 > $o/r02_bench_i.json 2> $o/r02_bench_i.err; echo "bench rc=$?"; cut -c1-260 $o/r02_bench_i.json; tail -3 $o/r02_bench_i.err
python profiles/c5_interp_sweep.py > $o/r02_c5_sweep_i.jsonl 2> $o/r02_c5_sweep_i.err; tail -4 $o/r02_c5_sweep_i.jsonl | cut -c1-400
