#!/bin/bash
# on the GPU box: one ncu --set full capture of the second vi_unit_kernel launch of tools/ncu_target.py
python tools/ncu_target.py 16 > gpurun_out/ncu_target.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:vi_unit_kernel -s 1 -c 1 -o gpurun_out/prof -f python tools/ncu_target.py 16 > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
# the streaming ingest kernels (third launch of each)
ncu --set full --clock-control none --import-source on -k regex:ingest_ -s 4 -c 2 -o gpurun_out/prof_ingest -f python tools/ncu_target_ingest.py 16 > gpurun_out/ncu_ingest.log 2>&1
tail -2 gpurun_out/ncu_ingest.log
