#!/bin/bash
# Round 2, GPU call 11 (one B200): device-resident Arnoldi loop + single-launch resistance-face update: full suite, headline bench.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider ) > gpurun_out/r02k_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02k_pytest.log
tail -5 gpurun_out/r02k_pytest.log
( time timeout 600 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
grep "^{" gpurun_out/r02k_bench.json | head -c 600; echo; tail -3 gpurun_out/r02k_bench.err
