#!/bin/bash
O=gpurun_out
for c in c4 c5; do
timeout 900 ncu --metrics gpu__time_duration.sum,sm__cycles_active.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file $O/r2_launches_$c.csv python tools/profile_step.py --config $c --throughput-tiles > $O/r2_profile_step_$c.log 2>&1; tail -1 $O/r2_profile_step_$c.log
done
