#!/bin/bash
# on the GPU box: full parity tests, phase profile (A/B of the threshold variants), short bench
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/t.log
VI_THRESHOLD=band python tools/phase_profile.py 64 > gpurun_out/phase64_band.txt 2>&1
python tools/phase_profile.py 64 > gpurun_out/phase64.txt 2>&1
python bench.py --no-cpu --no-ingest --no-extra --steps 30 > gpurun_out/bench_quick.json 2> gpurun_out/bench.err
cat gpurun_out/t.log; grep -E "kernel|threshold|thr:" gpurun_out/phase64_band.txt; cat gpurun_out/phase64.txt; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print('units/s', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])"
