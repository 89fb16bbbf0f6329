#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 120 python tools/attention_trace.py 512 16 1 | tee $O/r2_att_trace_tpc1.log
timeout 120 python tools/attention_trace.py 512 16 2 | tee $O/r2_att_trace_tpc2.log
timeout 120 python tools/attention_trace.py 2048 16 1 | head -3
timeout 120 python tools/attention_trace.py 2048 16 2 | head -3
B="python bench.py --steps 30 --warmup 3 --no-c3 --no-cpu-baseline --no-gpu-reference --no-roofline"
for rep in 1 2; do
for cfg in "tiles2:PSAM_GEMM_VARIANT=0x10:" "tiles1:PSAM_GEMM_VARIANT=0x10 PSAM_ATTENTION_TILES=1:" "twopass:PSAM_GEMM_VARIANT=0x10 PSAM_ATTENTION_TWOPASS=1:"; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}
  env $envs timeout 400 $B > $O/r2_ab5_${name}_$rep.json 2> $O/r2_ab5_${name}_$rep.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2_ab5_${name}_$rep.json").read().strip().splitlines()[-1])
    print("$rep $name", round(d["value"],1), "clouds/s  e2e", round(d["e2e"]["value"],1), " single-stream ms", round(d["run"]["single_stream_ms_per_cloud"],3), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name FAILED", e); print(open("$O/r2_ab5_${name}_$rep.err").read()[-800:])
PY
done
done
