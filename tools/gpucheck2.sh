#!/bin/bash
# on the GPU box: parity tests, phase profile, short bench, then one ncu full capture with source (development helper)
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t.log
python tools/phase_profile.py 64 > gpurun_out/phase64.txt 2>&1
python bench.py --no-cpu --no-ingest --steps 30 > gpurun_out/bench_quick.json 2> gpurun_out/bench.err
cat gpurun_out/t.log gpurun_out/phase64.txt; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print('units/s', d['value'], 'e2e', d['e2e'], 'frac', d['roofline']['frac'])"
VI_HOST_UPLOAD=copy python bench.py --no-cpu --no-ingest --steps 30 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('staged-upload e2e', d['e2e'])"
bash tools/gpuprof_unit.sh
