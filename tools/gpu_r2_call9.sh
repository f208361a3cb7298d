#!/bin/bash
# Round 2, GPU call 10 (TWO B200s): the fused product + exchange kernel: parity on all three transports, N = 2 bench fused vs unfused.
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 900 -p no:cacheprovider ) > gpurun_out/r02i_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02i_pytest.log
tail -6 gpurun_out/r02i_pytest.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 3 --warmup 3 --no-weak ) > gpurun_out/r02i_bench_n2_fused.json 2> gpurun_out/r02i_bench_n2_fused.err
( time SVB200_FUSED=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 3 --warmup 3 --no-weak ) > gpurun_out/r02i_bench_n2_unfused.json 2> gpurun_out/r02i_bench_n2_unfused.err
for f in fused unfused; do grep "^{" gpurun_out/r02i_bench_n2_$f.json | head -c 500; echo; tail -2 gpurun_out/r02i_bench_n2_$f.err; done
