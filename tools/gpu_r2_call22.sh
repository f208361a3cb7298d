#!/bin/bash
# Round 2, GPU call 22 (TWO B200s): the final tree: the whole GPU suite including the 2-rank multi-GPU cases.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider ) > gpurun_out/r02v_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02v_pytest.log
tail -6 gpurun_out/r02v_pytest.log
