#!/bin/bash
# Round 2, GPU call 12 (one B200): device-resident Arnoldi loop with the shared-memory Givens kernel: parity tests + headline bench.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dropin.py -m gpu -q --timeout 600 -p no:cacheprovider ) > gpurun_out/r02l_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02l_pytest.log
tail -5 gpurun_out/r02l_pytest.log
( time timeout 600 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err
grep "^{" gpurun_out/r02l_bench.json | head -c 400; echo; tail -3 gpurun_out/r02l_bench.err
