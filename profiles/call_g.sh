#!/bin/bash
o=gpurun_out
T="tests/test_gpu_path.py::test_training_and_rollout_loops_golden_fixture"
echo "--- TC interpolation"; timeout 300 python -m pytest $T -q -m gpu -x 2>&1 | grep -E "tensor\(\[|passed|failed" | head -3
echo "--- direct fp32 interpolation"; MMPDE_ITP_TC=0 timeout 300 python -m pytest $T -q -m gpu -x 2>&1 | grep -E "tensor\(\[|passed|failed" | head -3
echo "--- fp32 res_cut"; MMPDE_FP32_RES_CUT=1 timeout 300 python -m pytest $T -q -m gpu -x 2>&1 | grep -E "tensor\(\[|passed|failed" | head -3
echo "--- both"; MMPDE_ITP_TC=0 MMPDE_FP32_RES_CUT=1 timeout 300 python -m pytest $T -q -m gpu -x 2>&1 | grep -E "tensor\(\[|passed|failed" | head -3
make -C mm-pde_b200/csrc timeline > /dev/null 2>&1; python profiles/timeline_node.py > $o/r02_timeline_node_img.txt 2>&1; cat $o/r02_timeline_node_img.txt
