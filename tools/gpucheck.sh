#!/bin/bash
# on the GPU box: parity tests, phase profile, short bench (development helper; run through tools/gpu.sh)
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/t.log
python tools/phase_profile.py 64 > gpurun_out/phase64.txt 2>&1
python bench.py --no-cpu --steps 30 > gpurun_out/bench_quick.json 2> gpurun_out/bench.err
cat gpurun_out/t.log gpurun_out/phase64.txt; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print('units/s', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])"
