#!/bin/bash
# on the GPU box: A/B of the two asynchronous crop gathers (tensor boxes vs one bulk copy per row), bench only
for i in 1 2; do
python bench.py --no-cpu --no-ingest --no-extra --steps 40 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('tensor boxes: units/s', round(d['value']), 'ms', d['ms_per_step'])"
VI_GATHER=rows python bench.py --no-cpu --no-ingest --no-extra --steps 40 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('row copies  : units/s', round(d['value']), 'ms', d['ms_per_step'])"
done
