#!/bin/bash
# Round 2, GPU call 15 (one B200): two-stage multi-dot: parity tests, stand-alone tour of the Arnoldi kernels, headline bench.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider ) > gpurun_out/r02o_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02o_pytest.log
tail -4 gpurun_out/r02o_pytest.log
timeout 300 python tools/prof.py tour --reps 20 > gpurun_out/r02o_tour.jsonl 2> gpurun_out/r02o_tour.err
grep "multi_dot\|cgs" gpurun_out/r02o_tour.jsonl
( time timeout 600 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err
grep "^{" gpurun_out/r02o_bench.json | head -c 400; echo; tail -3 gpurun_out/r02o_bench.err
