#!/bin/bash
# single-GPU evidence for the committed build: full GPU suite, smoke, bench (default flags = what the driver runs), ncu
o=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > $o/r02_pytest_gpu_e.log 2>&1; echo "suite rc=$?"; tail -3 $o/r02_pytest_gpu_e.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > $o/r02_bench_g_final.json 2> $o/r02_bench_g_final.err; echo "bench rc=$?"; cut -c1-300 $o/r02_bench_g_final.json
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > $o/r02_bench_reference_arm.json 2> $o/r02_bench_reference_arm.err; echo "ref rc=$?"; cut -c1-400 $o/r02_bench_reference_arm.json
bash profiles/ncu_capture.sh r02b > $o/r02b_ncu_capture.log 2>&1; tail -15 $o/r02b_ncu_capture.log
