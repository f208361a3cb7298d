#!/bin/bash
# Round 2, GPU call 27 (one B200, ~40 s): the benchmark-size golden test on the very last build.
mkdir -p gpurun_out
( time timeout 55 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 50 -p no:cacheprovider -k "benchmark_size" -s ) > gpurun_out/r02zz_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02zz_pytest.log
grep -n "P10 counts\|passed\|failed\|rc" gpurun_out/r02zz_pytest.log
