#!/bin/bash
# next GPU call: tcgen05.cp probe, weight-image tests, full GPU suite, bench, cylinder breakdown
o=gpurun_out
./profiles/experiments/utccp_probe > $o/r02_utccp_probe.txt 2>&1; echo "probe rc=$?"; cat $o/r02_utccp_probe.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "weight_images or node_gemm" > $o/r02_pytest_e_img.log 2>&1; echo "img tests rc=$?"; tail -3 $o/r02_pytest_e_img.log
timeout 1500 python -m pytest tests -q -m gpu -x > $o/r02_pytest_e.log 2>&1; echo "suite rc=$?"; tail -3 $o/r02_pytest_e.log
python bench.py --steps 10 --warmup 3 > $o/r02_bench_e.json 2> $o/r02_bench_e.err; echo "bench rc=$?"; cut -c1-300 $o/r02_bench_e.json
python profiles/kineto_breakdown.py 3 --eager --ops --cylinder > $o/r02_kineto_cylinder_eager.txt 2>&1; echo "kineto rc=$?"; head -40 $o/r02_kineto_cylinder_eager.txt
