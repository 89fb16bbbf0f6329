#!/bin/bash
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 10 --warmup 3 --no-c3 --no-cpu-baseline --no-gpu-reference --no-roofline"
for cfg in "c2b8 2" "c2b8 3" "c2b8 4" "c2b4 4" "c2b4 6"; do
  set -- $cfg
  timeout 400 $B --config $1 --depth $2 > $O/r2_batch_$1_$2.json 2> $O/r2_batch_$1_$2.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2_batch_$1_$2.json").read().strip().splitlines()[-1])
    print("$1 depth $2:", round(d["value"],1), "clouds/s  e2e", round(d["e2e"]["value"],1), " single-stream ms/request", round(d["run"]["single_stream_ms_per_cloud"],3), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("$1 $2 FAILED", e); print(open("$O/r2_batch_$1_$2.err").read()[-600:])
PY
done
