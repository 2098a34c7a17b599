#!/bin/bash
# on the GPU box: full parity tests (with timings), phase profile, short bench, one ncu full capture
python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 > gpurun_out/t.log
python tools/phase_profile.py 64 > gpurun_out/phase64.txt 2>&1
python bench.py --no-cpu --steps 30 > gpurun_out/bench_quick.json 2> gpurun_out/bench.err
cat gpurun_out/t.log gpurun_out/phase64.txt; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print('units/s', d['value'], 'e2e', d['e2e']['value'], d['e2e_packed_masks']['value'], d['e2e_records_only']['value'], 'frac', d['roofline']['frac'])"
bash tools/gpuprof_unit.sh
