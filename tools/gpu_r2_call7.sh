#!/bin/bash
# Round 2, GPU call 7 (one B200): drop-in harness tests after the stdio fix, the FSI bench line with a valid state, ustruct line.
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_dropin.py tests/test_reference_main.py -m gpu -q --timeout 600 -p no:cacheprovider ) > gpurun_out/r02g_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02g_pytest.log
tail -4 gpurun_out/r02g_pytest.log
for w in fsi_pipe ustruct_block; do
  ( time timeout 900 python bench.py --workload $w --steps 3 --warmup 2 ) > gpurun_out/r02g_bench_$w.json 2> gpurun_out/r02g_bench_$w.err
  head -c 600 gpurun_out/r02g_bench_$w.json; echo; tail -2 gpurun_out/r02g_bench_$w.err
done
