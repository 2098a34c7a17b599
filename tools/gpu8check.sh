#!/bin/bash
# on an 8-GPU box: configs[2] as stated (1,024 frames sharded by image) and the host-link ceiling with every rank copying at once
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tools/host_link_probe.py --bind 1 > gpurun_out/probe_n8.txt 2> gpurun_out/probe_n8.err
tail -3 gpurun_out/bench_n8.err; head -c 1200 gpurun_out/bench_n8.json; echo; grep probe gpurun_out/probe_n8.txt
