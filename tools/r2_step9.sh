#!/bin/bash
O=gpurun_out
mkdir -p $O
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee $O/r2_tests9.log
grep -q " failed\| error" $O/r2_tests9.log && exit 1
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-roofline"
for cfg in "default::" ; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}
  env $envs timeout 500 $B > $O/r2_ab7_${name}.json 2> $O/r2_ab7_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2_ab7_${name}.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"],1), "clouds/s  e2e", round(d["e2e"]["value"],1), " c3", round(d["c3"]["value"],1), round(d["c3"]["e2e"]["value"],1), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name FAILED", e); print(open("$O/r2_ab7_${name}.err").read()[-800:])
PY
done
