#!/bin/bash
# on a 2-GPU box: the two-rank parity test, a short configs[2]-style bench at N=2, the host-link probe
nvidia-smi topo -m > gpurun_out/topo2.txt 2>&1
python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --global-images 256 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
for b in 0 1; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/host_link_probe.py --bind $b >> gpurun_out/probe_n2.txt 2>> gpurun_out/probe_n2.err
done
python tools/host_link_probe.py --bind 1 >> gpurun_out/probe_n2.txt 2>> gpurun_out/probe_n2.err
cat gpurun_out/t2.log; tail -5 gpurun_out/bench_n2.err; head -c 1500 gpurun_out/bench_n2.json; echo; cat gpurun_out/probe_n2.txt
