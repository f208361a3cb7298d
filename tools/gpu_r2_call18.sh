#!/bin/bash
# Round 2, GPU call 18 (one B200): the final tree: full GPU suite, smoke, headline bench (with cpu_baseline / same_config), reference arm.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider ) > gpurun_out/r02r_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02r_pytest.log
tail -5 gpurun_out/r02r_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02r_smoke.log 2>&1; tail -2 gpurun_out/r02r_smoke.log
( time timeout 600 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02r_bench_ref.json 2> gpurun_out/r02r_bench_ref.err
grep "^{" gpurun_out/r02r_bench.json | head -c 400; echo; grep "^{" gpurun_out/r02r_bench_ref.json | head -c 400; echo
