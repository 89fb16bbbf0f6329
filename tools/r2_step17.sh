#!/bin/bash
O=gpurun_out
echo "== attention tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" 2>&1 | tail -8 | tee $O/r2_tests17.log
grep -q " failed\| error" $O/r2_tests17.log && exit 1
echo "== c5 parity"; timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "config5 or golden" 2>&1 | tail -5 | tee $O/r2_tests17b.log
grep -q " failed\| error" $O/r2_tests17b.log && exit 1
for cfg in "fused88::" "unfused88:PSAM_FUSED_ATTENTION_DH88=0:"; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}
  env $envs python bench.py --config c5 --depth 4 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-roofline > $O/r2_c5_$name.json 2> $O/r2_c5_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2_c5_$name.json").read().strip().splitlines()[-1])
    print("$name c5", round(d["value"],1), "clouds/s  e2e", round(d["e2e"]["value"],1), "launches", d["launches_per_cloud"])
except Exception as ex:
    print("$name FAILED", ex, open("$O/r2_c5_$name.err").read()[-600:])
PY
done
