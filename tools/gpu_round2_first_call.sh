#!/bin/bash
# First GPU call of round 2 (one B200, ~12 minutes): everything that could not be re-measured after the last session of round 1.
#   gpurun --timeout 900 -- 'bash tools/gpu_round2_first_call.sh'
#   1. full GPU suite with the pending marks lifted (--runxfail): the six pending tests of tests/test_zz_late_additions.py (FSI wall, result-file comparison, solid block end to end) and the
#      three P10-size property tests are the ones that have not run since the last additions
#   2. smoke(), the bench line (own arm + reference arm)
#   3. event timings + one ncu capture of the extended struct / ustruct element (VISC = true instantiations: 252 registers; the
#      ustruct one spills 800 bytes of stack)
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q --runxfail --timeout 300 -p no:cacheprovider ) > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02a_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_smoke.log 2>&1
( time timeout 600 python bench.py --steps 3 --warmup 3 ) > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
( time timeout 300 python bench.py --impl reference --steps 1 --warmup 1 ) > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_assemble_solid -c 6 -o gpurun_out/r02a_solid_ext \
    python -m pytest tests/test_zz_late_additions.py -m gpu -q --runxfail -k "viscosity and hex" -p no:cacheprovider > gpurun_out/r02a_ncu.log 2>&1
bash tools/ncu_export.sh gpurun_out/r02a_solid_ext
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/r02a_smi.txt
tail -4 gpurun_out/r02a_pytest.log; tail -2 gpurun_out/r02a_smoke.log; head -c 700 gpurun_out/r02a_bench.json; echo; head -c 300 gpurun_out/r02a_bench_ref.json

# Second call (two GPUs, ~6 minutes): where the N = 2 line loses its 28 % - kernel classes incl. "halo" and the all-reduce share are in
# the bench line's kernel_shares / kernels objects, NVLS use in the NCCL log.
#   gpurun --gpus 2 --timeout 600 -- 'NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
#       --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02b_bench_n2.json 2> gpurun_out/r02b_bench_n2.err; \
#       python -m pytest tests/test_multigpu.py -m gpu -q > gpurun_out/r02b_multigpu.log 2>&1; tail -3 gpurun_out/r02b_multigpu.log'
