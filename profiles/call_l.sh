#!/bin/bash
for i in 1 2 3; do timeout 300 python -m pytest "tests/test_gpu_path.py::test_partitioned_solver_equals_whole_graph" -q -m gpu 2>&1 | grep -E "^E  |passed|failed" | head -8; done
timeout 900 python -m pytest tests/test_gpu_path.py tests/test_gpu_properties.py tests/test_gpu_fullsize.py -q -m gpu 2>&1 | tail -5
