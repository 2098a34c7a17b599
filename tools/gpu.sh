#!/bin/bash
# build the extension here, then run the given command on the GPU box (development helper)
set -e
cd /root/repo
python -c "import __graft_entry__ as g; g.build()"
exec /usr/local/graft/bin/gpurun --timeout ${GPU_TIMEOUT:-900} -- "$@"
