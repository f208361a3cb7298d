#!/bin/bash
# Round 2, GPU call 8 (FOUR B200s): multi-GPU parity incl. the 4-rank cases (z-slabs, RCB, METIS partitions), the N = 4 bench line
# (strong headline + weak object), A/B of the vv3 variants incl. the 32-register one.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02h_topo.txt 2>&1
( time timeout 1500 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 900 -p no:cacheprovider ) > gpurun_out/r02h_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02h_pytest.log
tail -6 gpurun_out/r02h_pytest.log
( time NCCL_DEBUG=INFO timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 3 --warmup 3 ) > gpurun_out/r02h_bench_n4.json 2> gpurun_out/r02h_bench_n4.err
grep -v "^{" gpurun_out/r02h_bench_n4.json | head -5; grep "^{" gpurun_out/r02h_bench_n4.json | head -c 1500; echo
grep -i "nvls\|via P2P\|Connected all" gpurun_out/r02h_bench_n4.err | head -5
timeout 300 python tools/prof.py tiled --reps 20 > gpurun_out/r02h_variants.jsonl 2> gpurun_out/r02h_variants.err
head -8 gpurun_out/r02h_variants.jsonl
