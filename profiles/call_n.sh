#!/bin/bash
# edge kernels: kernel / layer / solver / guard tests, then stand-alone timing
o=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_guards.py -q -m gpu -x -k "edge or layer or solver" > $o/r02_pytest_edge.log 2>&1; echo "tests rc=$?"; tail -3 $o/r02_pytest_edge.log
timeout 300 python profiles/edge_bench.py 30 2>&1 | grep edge_
