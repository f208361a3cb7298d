#!/bin/bash
# Reduce an .ncu-rep on the GPU box to the CSV pages we read here (gpurun_out/ is capped at 64 MiB):
#   tools/ncu_export.sh gpurun_out/<name>   ->  <name>_raw.csv (+ <name>_source.csv when small), rep deleted if > 24 MiB
set -e
rep="$1.ncu-rep"
ncu -i "$rep" --page raw --csv > "$1_raw.csv" 2>/dev/null || true
ncu -i "$rep" --page details --csv > "$1_details.csv" 2>/dev/null || true
sz=$(stat -c %s "$rep")
if [ "$sz" -gt 25165824 ]; then rm -f "$rep"; echo "removed $rep ($sz bytes)"; fi
