#!/bin/bash
# on the GPU box: the host-buffer call's chunk size (images per chunk) against its end-to-end rate
for c in 2 3 4 5 7 10; do
VI_HOST_CHUNK=$c python bench.py --no-cpu --no-ingest --no-extra --steps 12 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunk $c: e2e', round(d['e2e']['value']), 'units/s', d['e2e']['ms_per_step'], 'ms')"
done
