#!/bin/bash
# compute-sanitizer passes over the hand-rolled mbarrier / TMEM pipelines at small sizes (run on the GPU box):
#   memcheck  : out-of-bounds / misaligned global + shared accesses in every kernel the selected tests launch
#   racecheck : shared-memory hazards between the warp roles (the operand tiles are handed over through mbarriers)
# usage: bash profiles/sanitize.sh <output prefix>      -> <prefix>_memcheck.log, <prefix>_racecheck.log
out=${1:-gpurun_out/r02_sanitizer}
SEL='test_edge_tensor_core_kernels_match_fp64_definition and (sizes0 or sizes3) or test_edge_kernels_empty_and_single_edge or test_node_gemm_tensor_core and (-300 or -127) or test_node_wgrad_tensor_core and (63 or 1000) or test_node_wgrad_grouped and sizes2 or test_knn_grid_multi_equals_single_searches or test_tensor_core_interpolation_equals_direct_form and (129 or 64-7) or test_batchnorm_forward_backward_and_running_stats'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
      python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "$SEL" > ${out}_$tool.log 2>&1
  echo "$tool rc=$?" >> ${out}_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" ${out}_$tool.log | tail -5
done
