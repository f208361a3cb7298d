#!/bin/bash
# Round 2, GPU call 1 (one B200): full GPU suite with the pending marks removed, smoke, the default bench line.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider ) > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02a_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_smoke.log 2>&1
( time timeout 600 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/r02a_smi.txt
nvidia-smi topo -m > gpurun_out/r02a_topo.txt 2>&1
tail -6 gpurun_out/r02a_pytest.log; tail -2 gpurun_out/r02a_smoke.log; head -c 600 gpurun_out/r02a_bench.json
