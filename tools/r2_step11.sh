#!/bin/bash
O=gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-c3 --no-cpu-baseline --no-gpu-reference --no-roofline --depth 4"
for cfg in "tp_fold::" "lat_fold:PSAM_THROUGHPUT_TILES=0:" "tp_nofold:PSAM_FUSED_BLOCK_LN=0:" "lat_nofold:PSAM_THROUGHPUT_TILES=0 PSAM_FUSED_BLOCK_LN=0:"; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}
  env $envs timeout 300 $B > $O/r2_d4_${name}.json 2> $O/r2_d4_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2_d4_${name}.json").read().strip().splitlines()[-1])
    print("depth4 $name", round(d["value"],1), "clouds/s  e2e", round(d["e2e"]["value"],1), " single-stream ms", round(d["run"]["single_stream_ms_per_cloud"],3))
except Exception as e:
    print("$name FAILED", e); print(open("$O/r2_d4_${name}.err").read()[-600:])
PY
done
