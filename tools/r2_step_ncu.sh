#!/bin/bash
# ncu captures; the .ncu-rep stay on the box (/tmp), only CSV pages come back (gpurun_out/ is capped at 64 MiB).
O=gpurun_out
mkdir -p $O
for k in "$@"; do
  timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -o /tmp/r2_$k -f python tools/kernel_once.py $k > $O/r2_ncu_$k.log 2>&1
  tail -2 $O/r2_ncu_$k.log
  ncu -i /tmp/r2_$k.ncu-rep --page raw --csv > $O/r2_$k.raw.csv 2>/dev/null
  ncu -i /tmp/r2_$k.ncu-rep --page source --csv > $O/r2_$k.source.csv 2>/dev/null
  ls -la /tmp/r2_$k.ncu-rep $O/r2_$k.raw.csv $O/r2_$k.source.csv
done
