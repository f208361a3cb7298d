#!/bin/bash
# Round 2, GPU call 25 (one B200, short): the plain GMRES variant of SURVEY 8(d) at P10 with its reference golden.
mkdir -p gpurun_out
( time timeout 140 python bench.py --ls GMRES --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02y_bench_gmres.json 2> gpurun_out/r02y_bench_gmres.err
grep "^{" gpurun_out/r02y_bench_gmres.json | head -c 500; echo; tail -3 gpurun_out/r02y_bench_gmres.err
