#!/bin/bash
# Round 2, GPU call 21 (one B200): component-wise copy of G for Schur pass 1: variants parity test, stand-alone A/B, in-step A/B.
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider -k "tma_staged" ) > gpurun_out/r02u_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02u_pytest.log; tail -3 gpurun_out/r02u_pytest.log
timeout 300 python tools/prof.py tiled --reps 20 > gpurun_out/r02u_variants.jsonl 2> gpurun_out/r02u_variants.err; grep spmv_sv gpurun_out/r02u_variants.jsonl
for v in 0 3; do
  ( SVB200_SCHUR_GP=$v timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02u_bench_gp$v.json 2> gpurun_out/r02u_bench_gp$v.err
  python - $v <<'PY'
import json, sys
v = sys.argv[1]
d = json.loads([l for l in open(f"gpurun_out/r02u_bench_gp{v}.json").read().splitlines() if l.startswith("{")][-1])
print("schur_gp", v, "ms", round(d["ms_per_step"], 1), "spmv_sv", d["kernels"]["spmv_sv"], "counts", d["run"]["krylov_itr"], d["run"]["gm_itr"], d["run"]["cg_itr"])
PY
done
