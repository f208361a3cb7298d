#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -X faulthandler -m pytest tests/test_dropin.py -m gpu -q --timeout 600 -p no:cacheprovider -x -s ) > gpurun_out/r02f_dropin.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02f_dropin.log
tail -40 gpurun_out/r02f_dropin.log
( time timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider --deselect tests/test_dropin.py ) > gpurun_out/r02f_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02f_pytest.log
tail -5 gpurun_out/r02f_pytest.log
