#!/bin/bash
# Round 2, GPU call 14 (EIGHT B200s): the N = 8 bench line: 10M-tet pipe split over 8 ranks (strong headline), P80 (weak object).
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02n_topo.txt 2>&1
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 3 --warmup 3 ) > gpurun_out/r02n_bench_n8.json 2> gpurun_out/r02n_bench_n8.err
grep "^{" gpurun_out/r02n_bench_n8.json | head -c 1200; echo; tail -3 gpurun_out/r02n_bench_n8.err
