#!/bin/bash
# Round 2, GPU call 2 (one B200): the new parity tests (plug-in inside the reference's main(), P10 golden, coupled time step,
# TMA-staged SpMV variants), A/B of the tiled kernels at P10 and one ncu capture of them.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_reference_main.py tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider -s ) > gpurun_out/r02b_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02b_pytest.log
timeout 600 python tools/prof.py tiled --reps 20 > gpurun_out/r02b_tiled.jsonl 2> gpurun_out/r02b_tiled.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_tiled -c 6 -o gpurun_out/r02b_tiled_ncu \
    python tools/prof.py tiled --reps 1 > gpurun_out/r02b_ncu.log 2>&1
bash tools/ncu_export.sh gpurun_out/r02b_tiled_ncu >> gpurun_out/r02b_ncu.log 2>&1
tail -5 gpurun_out/r02b_pytest.log; cat gpurun_out/r02b_tiled.jsonl; tail -3 gpurun_out/r02b_tiled.err
