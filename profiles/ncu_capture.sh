#!/bin/bash
# Run on the GPU box (gpurun): evidence for the committed build.
#   1. plain bench (must exit 0 before anything runs under ncu)
#   2. ncu launch list of the same command (gpu__time_duration only): shares per kernel
#   3. one `ncu --set full` capture per hot kernel (a warm launch of the second eager step)
# usage: bash profiles/ncu_capture.sh <tag>        -> gpurun_out/<tag>_*
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
BENCH="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
$BENCH > $out/${tag}_bench_for_ncu.json 2> $out/${tag}_bench_for_ncu.err || { echo "bench failed"; tail -5 $out/${tag}_bench_for_ncu.err; exit 1; }
tail -c 400 $out/${tag}_bench_for_ncu.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/${tag}_launches.csv \
    $BENCH > $out/${tag}_launches_run.log 2>&1
echo "launch list rc=$?"
EAGER="python bench.py --steps 1 --warmup 1 --no-graph --no-extras --no-cpu-baseline"
# kernel regex : launches of that kernel to skip (lands in the second step, middle layer)
for spec in edge_bwd_tc:17 edge_fwd_tc:15 node_gemm_tc:130 node_wgrad_tc:22 itp_tc:6 decoder_fwd:2 \
            bn_stats:20 bn_apply_kernel:20 bn_bwd_reduce:20 bn_bwd_apply:20 knn_grid_kernel:1; do
  k=${spec%%:*}; s=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o $out/${tag}_full_$k \
      $EAGER > $out/${tag}_full_$k.log 2>&1
  echo "$k rc=$?"
done
ls -la $out/${tag}_full_*.ncu-rep
