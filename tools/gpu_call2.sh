#!/bin/bash
# full validation: GPU suite, smoke, bench (own arm + reference arm), ncu launch list of the bench command
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --timeout 300 ) > gpurun_out/r01m_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r01m_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01m_smoke.log 2>&1
( time timeout 900 python bench.py --steps 3 --warmup 3 ) > gpurun_out/r01m_bench.json 2> gpurun_out/r01m_bench.err
( time timeout 600 python bench.py --impl reference --steps 1 --warmup 1 ) > gpurun_out/r01m_bench_ref.json 2> gpurun_out/r01m_bench_ref.err
tail -4 gpurun_out/r01m_pytest.log; tail -2 gpurun_out/r01m_smoke.log; head -c 600 gpurun_out/r01m_bench.json; echo; tail -4 gpurun_out/r01m_bench.err; head -c 300 gpurun_out/r01m_bench_ref.json
