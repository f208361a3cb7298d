#!/bin/bash
# A/B at N = 1: device-resident Arnoldi loop on / off, face fusion on / off, batch size
mkdir -p gpurun_out
for v in "default" "SVB200_GMRES_DEVICE=0" "SVB200_FACE_FUSED=0" "SVB200_GM_BATCH=4" "SVB200_GM_BATCH=16"; do
  tag=$(echo $v | tr '=' '_')
  ( [ "$v" != "default" ] && export $v; timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02m_$tag.json 2> gpurun_out/r02m_$tag.err
  python - "$tag" <<'PY'
import json, sys
t = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r02m_{t}.json").read().splitlines() if l.startswith("{")][-1])
    print(t, "ms", round(d["ms_per_step"], 1), "kfrac", round(d["kernel_time_frac_of_step"], 3), "launches/step", d["gpu_launches"] / d["steps"], "fixed", round(d["fixed_work"]["ms_per_step"], 1))
except Exception as e:
    print(t, "failed", e)
PY
done
