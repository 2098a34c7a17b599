#!/bin/bash
# on the GPU box: phase profile and a short bench only (quick look between two full checks)
python tools/phase_profile.py 64 > gpurun_out/phase64.txt 2>&1
python bench.py --no-cpu --no-ingest --no-extra --steps 30 > gpurun_out/bench_quick.json 2> gpurun_out/bench.err
cat gpurun_out/phase64.txt; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print('units/s', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])"
