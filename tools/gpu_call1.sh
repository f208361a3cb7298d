#!/bin/bash
# round-1 session-4 GPU call: full GPU suite, generic fluid element timings + ncu capture
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q --timeout 300 ) > gpurun_out/r01e_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r01e_pytest.log
timeout 400 python tools/prof.py fluidgen --reps 5 > gpurun_out/r01e_fluidgen_events.jsonl 2> gpurun_out/r01e_fluidgen.err
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_assemble_fluid_gen -c 4 -o gpurun_out/r01e_fluidgen python tools/prof.py fluidgen --reps 1 > gpurun_out/r01e_ncu.log 2>&1
bash tools/ncu_export.sh gpurun_out/r01e_fluidgen
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r01e_smi.txt
tail -5 gpurun_out/r01e_pytest.log; cat gpurun_out/r01e_fluidgen_events.jsonl
