#!/bin/bash
# Round 2, GPU call 24 (TWO B200s, short): the host tables moved to lhs_layout.hpp (halo source lists, row tiles) on the device paths.
mkdir -p gpurun_out
( time timeout 150 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 140 -p no:cacheprovider -k "test_partitioned_solve_matches_reference and 2-NS-p2p_fused" ) > gpurun_out/r02x_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02x_pytest.log
( time timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 90 -p no:cacheprovider -k "tma_staged and dims0 or test_solve_matches_golden" ) >> gpurun_out/r02x_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02x_pytest.log
grep -n "passed\|failed\|rc" gpurun_out/r02x_pytest.log
