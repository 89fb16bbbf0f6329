#!/bin/bash
O=gpurun_out
mkdir -p $O
echo "== attention tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "fused_attention" 2>&1 | tail -25 | tee $O/r2_att2.log
grep -q "failed\|error" $O/r2_att2.log && exit 1
bash tools/r2_step_ncu.sh attention
B="python bench.py --steps 30 --warmup 3 --no-c3 --no-cpu-baseline --no-gpu-reference --no-roofline"
for cfg in "fold_nodual:PSAM_GEMM_VARIANT=0x10:" "nofold_dual:PSAM_FUSED_BLOCK_LN=0:"; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}
  env $envs timeout 400 $B > $O/r2_ab4_${name}.json 2> $O/r2_ab4_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2_ab4_${name}.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"],1), "clouds/s  e2e", round(d["e2e"]["value"],1), " single-stream ms", round(d["run"]["single_stream_ms_per_cloud"],3), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name FAILED", e); print(open("$O/r2_ab4_${name}.err").read()[-800:])
PY
done
