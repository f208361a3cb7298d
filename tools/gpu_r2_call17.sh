#!/bin/bash
# Round 2, GPU call 17 (TWO B200s): one-launch face update + Givens riding on the update kernel: parity (1 GPU and 2 GPUs), bench N = 1, 2.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dropin.py -m gpu -q --timeout 600 -p no:cacheprovider ) > gpurun_out/r02q_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02q_pytest.log
tail -4 gpurun_out/r02q_pytest.log
( time timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 600 -p no:cacheprovider -k "p2p_fused or slab" ) > gpurun_out/r02q_pytest2.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02q_pytest2.log
tail -4 gpurun_out/r02q_pytest2.log
( time timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02q_bench_n1.json 2> gpurun_out/r02q_bench_n1.err
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 2 --steps 3 --warmup 3 --no-weak ) > gpurun_out/r02q_bench_n2.json 2> gpurun_out/r02q_bench_n2.err
for n in 1 2; do grep "^{" gpurun_out/r02q_bench_n$n.json | head -c 300; echo; done
