#!/bin/bash
o=gpurun_out
T="tests/test_gpu_path.py::test_training_and_rollout_loops_golden_fixture"
for i in 1 2; do timeout 300 python -m pytest $T -q -m gpu -x 2>&1 | grep -E "tensor\(\[|passed|failed" | head -4; done
echo "--- images off"
for i in 1 2; do MMPDE_WEIGHT_IMAGES=0 timeout 300 python -m pytest $T -q -m gpu -x 2>&1 | grep -E "tensor\(\[|passed|failed" | head -4; done
echo "--- rest of the suite"
timeout 1500 python -m pytest tests -q -m gpu --deselect $T > $o/r02_pytest_f.log 2>&1; echo "suite rc=$?"; tail -8 $o/r02_pytest_f.log
