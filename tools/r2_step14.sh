#!/bin/bash
O=gpurun_out
echo "== rowln test"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "rowln" 2>&1 | tail -12 | tee $O/r2_tests14a.log
grep -q " failed\| error" $O/r2_tests14a.log && exit 1
echo "== model tests"; timeout 1500 python -m pytest tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -8 | tee $O/r2_tests14b.log
grep -q " failed\| error" $O/r2_tests14b.log && exit 1
for cfg in "fused::" "unfused:PSAM_FUSED_ROW_LN=0:"; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}
  env $envs python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-roofline > $O/r2_ab10_$name.json 2> $O/r2_ab10_$name.err
  env $envs python bench.py --config c4 --depth 4 --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-roofline > $O/r2_ab10c4_$name.json 2> $O/r2_ab10c4_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2_ab10_$name.json").read().strip().splitlines()[-1])
    e=json.loads(open("$O/r2_ab10c4_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"],1), "clouds/s  e2e", round(d["e2e"]["value"],1), " c3", round(d["c3"]["value"],1), "clk", d["clocks"]["sm_mhz"], " c4", round(e["value"],1))
except Exception as ex:
    print("$name FAILED", ex, open("$O/r2_ab10_$name.err").read()[-600:])
PY
done
