#!/bin/bash
# Round 2, GPU call 19 (EIGHT B200s): the N = 8 line of the final tree (strong headline + weak object).
mkdir -p gpurun_out
( time NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29599 bench.py --gpus 8 --steps 3 --warmup 3 ) > gpurun_out/r02s_bench_n8.json 2> gpurun_out/r02s_bench_n8.err
grep "^{" gpurun_out/r02s_bench_n8.json | head -c 400; echo; tail -3 gpurun_out/r02s_bench_n8.err
