#!/bin/bash
# on the GPU box: quick experiment check (subset of parity tests, phase profile, short bench)
python -m pytest tests -m gpu -x -q -k "goldens or oracle_records or full_size" 2>&1 | tail -5 > gpurun_out/t.log
python tools/phase_profile.py 64 > gpurun_out/phase64.txt 2>&1
python bench.py --no-cpu --no-ingest --steps 30 > gpurun_out/bench_quick.json 2> gpurun_out/bench.err
cat gpurun_out/t.log gpurun_out/phase64.txt; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print('units/s', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])"
