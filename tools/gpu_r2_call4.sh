#!/bin/bash
# Round 2, GPU call 4 (one B200): vv3 variant A/B, the variants parity test, FP64 FMA peak, bench lines of configs[3] / configs[4].
mkdir -p gpurun_out
timeout 300 python tools/prof.py tiled --reps 20 > gpurun_out/r02d_variants.jsonl 2> gpurun_out/r02d_variants.err
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_reference_main.py -m gpu -q --timeout 600 -p no:cacheprovider -k "tma_staged or ustruct" ) > gpurun_out/r02d_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02d_pytest.log
python - > gpurun_out/r02d_fp64_peak.json 2> gpurun_out/r02d_fp64_peak.err <<'PY'
import json, sys
sys.path.insert(0, ".")
from svfsiplus_b200 import backend as B
be = B.Backend(0)
print(json.dumps({"fp64_fma_peak_tflops": [be.fp64_peak(k=k, reps=5) for k in (4, 16, 64)], "how": "k_fma_peak: 148 x 8 CTAs x 256 threads, 8 independent DFMA chains per thread, CUDA events"}))
PY
for w in struct_block ustruct_block fsi_pipe; do
  ( time timeout 900 python bench.py --workload $w --steps 3 --warmup 2 ) > gpurun_out/r02d_bench_$w.json 2> gpurun_out/r02d_bench_$w.err
done
cat gpurun_out/r02d_variants.jsonl | head -8; tail -3 gpurun_out/r02d_pytest.log; cat gpurun_out/r02d_fp64_peak.json; for w in struct_block ustruct_block fsi_pipe; do head -c 900 gpurun_out/r02d_bench_$w.json; echo; tail -2 gpurun_out/r02d_bench_$w.err; done
