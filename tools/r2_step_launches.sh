#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum,sm__cycles_active.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file $O/r2_launches_c2.csv python tools/profile_step.py --config c2 --throughput-tiles > $O/r2_profile_step.log 2>&1; tail -2 $O/r2_profile_step.log
ls -la $O/r2_launches_c2.csv
