#!/bin/bash
# Round 2, GPU call 16 (EIGHT B200s): final multi-GPU lines: N = 8, then N = 4 and N = 2 on the same box (subsets of its GPUs).
mkdir -p gpurun_out
for n in 8 4 2; do
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29580+n)) bench.py --gpus $n --steps 3 --warmup 3 $( [ $n -lt 8 ] && echo --no-weak ) ) > gpurun_out/r02p_bench_n$n.json 2> gpurun_out/r02p_bench_n$n.err
  grep "^{" gpurun_out/r02p_bench_n$n.json | head -c 300; echo
done
