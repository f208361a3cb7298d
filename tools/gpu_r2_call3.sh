#!/bin/bash
# Round 2, GPU call 3 (TWO B200s): (a) one-GPU leftovers: ustruct through main(), SpMV variants test, A/B of the vv3 variants;
# (b) multi-GPU parity on both transports (peer-mapped windows / NCCL); (c) the N = 2 bench line on both transports.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02c_topo.txt 2>&1
( time timeout 600 python -m pytest tests/test_reference_main.py tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider -k "ustruct or tma_staged" ) > gpurun_out/r02c_pytest1.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02c_pytest1.log
timeout 300 python tools/prof.py tiled --reps 20 > gpurun_out/r02c_variants.jsonl 2> gpurun_out/r02c_variants.err
( time timeout 1200 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 900 -p no:cacheprovider ) > gpurun_out/r02c_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02c_pytest.log
tail -5 gpurun_out/r02c_pytest1.log; cat gpurun_out/r02c_variants.jsonl; tail -5 gpurun_out/r02c_pytest.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/r02c_bench_n2_p2p.json 2> gpurun_out/r02c_bench_n2_p2p.err
( time SVB200_P2P=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/r02c_bench_n2_nccl.json 2> gpurun_out/r02c_bench_n2_nccl.err
head -c 1200 gpurun_out/r02c_bench_n2_p2p.json; echo; tail -3 gpurun_out/r02c_bench_n2_p2p.err
