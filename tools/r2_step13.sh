#!/bin/bash
O=gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "predictor or serving or eval_driver or iterative" 2>&1 | tail -8 | tee $O/r2_tests13.log
grep -q " failed\| error" $O/r2_tests13.log && exit 1
for rep in 1 2; do
python bench.py --steps 30 --warmup 3 --no-c3 --no-cpu-baseline --no-gpu-reference --no-roofline > $O/r2_ab9_$rep.json 2> $O/r2_ab9_$rep.err
python - <<PY
import json
try:
    d=json.loads(open("$O/r2_ab9_$rep.json").read().strip().splitlines()[-1])
    print("$rep", round(d["value"],1), "clouds/s  e2e", round(d["e2e"]["value"],1), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("FAILED", e, open("$O/r2_ab9_$rep.err").read()[-800:])
PY
done
