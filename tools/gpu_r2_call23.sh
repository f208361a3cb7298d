#!/bin/bash
# Round 2, GPU call 23 (one B200): the Givens column arithmetic shared between host and device: parity tests.
mkdir -p gpurun_out
( time timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -p no:cacheprovider ) > gpurun_out/r02w_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02w_pytest.log
tail -4 gpurun_out/r02w_pytest.log
