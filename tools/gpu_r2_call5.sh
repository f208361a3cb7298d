#!/bin/bash
# Round 2, GPU call 5 (one B200): full GPU suite with the column-owner vv3 kernel as default, the headline bench line, the
# configs[3] / configs[4] lines again (FSI state fixed), ncu capture of the vv3 variants.
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider ) > gpurun_out/r02e_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02e_pytest.log
tail -4 gpurun_out/r02e_pytest.log
( time timeout 600 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
for w in struct_block fsi_pipe; do
  ( time timeout 900 python bench.py --workload $w --steps 3 --warmup 2 ) > gpurun_out/r02e_bench_$w.json 2> gpurun_out/r02e_bench_$w.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_vv -c 8 -o gpurun_out/r02e_vv3_ncu \
    python tools/prof.py tiled --reps 1 > gpurun_out/r02e_ncu.log 2>&1
bash tools/ncu_export.sh gpurun_out/r02e_vv3_ncu >> gpurun_out/r02e_ncu.log 2>&1
head -c 700 gpurun_out/r02e_bench.json; echo; for w in struct_block fsi_pipe; do head -c 400 gpurun_out/r02e_bench_$w.json; echo; done
