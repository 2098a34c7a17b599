#!/bin/bash
# on the GPU box: one ncu --set full capture (with source) of the second vi_unit_kernel launch of tools/ncu_target.py
python tools/ncu_target.py 16 > gpurun_out/ncu_target.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:vi_unit_kernel -s 1 -c 1 -o gpurun_out/prof -f python tools/ncu_target.py 16 > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
