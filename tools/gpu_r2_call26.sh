#!/bin/bash
# Round 2, GPU call 26 (one B200, short): the headline line with the profile counters corrected for skipped iterations.
mkdir -p gpurun_out
( time timeout 110 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err
grep "^{" gpurun_out/r02z_bench.json | head -c 300; echo; tail -2 gpurun_out/r02z_bench.err
