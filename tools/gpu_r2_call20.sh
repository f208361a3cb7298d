#!/bin/bash
# Round 2, GPU call 20 (one B200): ncu launch list of ONE bench step (gpu__time_duration per launch) and a --set full capture of the
# step's dominant kernels taken from the middle of the step (deep Krylov basis).
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02t_step_launches.csv \
    python tools/step_launches.py > gpurun_out/r02t_step.log 2>&1
tail -2 gpurun_out/r02t_step.log
python tools/step_launches.py --summarise gpurun_out/r02t_step_launches.csv > gpurun_out/r02t_step_launches_summary.json 2> gpurun_out/r02t_sum.err
head -c 1500 gpurun_out/r02t_step_launches_summary.json
gzip -f gpurun_out/r02t_step_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_spmv_vv3c|k_schur_gp|k_schur_sp4|k_multi_dot|k_cgs_update_scale" \
    --launch-skip 6000 -c 14 -o gpurun_out/r02t_step_kernels python tools/step_launches.py > gpurun_out/r02t_ncu.log 2>&1
bash tools/ncu_export.sh gpurun_out/r02t_step_kernels >> gpurun_out/r02t_ncu.log 2>&1
tail -3 gpurun_out/r02t_ncu.log
