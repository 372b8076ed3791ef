#!/bin/bash
python profiles/debug/itp_ab.py 12 2>&1 | grep -v Warn | head -60
./profiles/experiments/launch_floor_probe 2>&1 | tee gpurun_out/r02_launch_floor_probe.txt
