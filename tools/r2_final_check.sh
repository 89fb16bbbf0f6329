#!/bin/bash
# Final health check of the committed state (tight time limits: the round's GPU budget is nearly spent).
O=gpurun_out
echo "== tests"; timeout 240 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee $O/r2_final_tests.log
echo "== bench"; timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference > $O/r2_final_bench.json 2> $O/r2_final_bench.err
python - <<PY
import json
try:
    d=json.loads(open("$O/r2_final_bench.json").read().strip().splitlines()[-1])
    print("bench", round(d["value"],1), "clouds/s  e2e", round(d["e2e"]["value"],1), " c3", round(d["c3"]["value"],1), "frac", round(d["roofline"]["frac"],3), "tok", round(d["tokenizer_in_regime"]["frac_of_hbm"],3), "knn", d["knn"]["ms"])
except Exception as ex:
    print("bench FAILED", ex, open("$O/r2_final_bench.err").read()[-600:])
PY
echo "== dh88 in the model (opt-in path, 60 s limit)"; PSAM_FUSED_ATTENTION_DH88=1 timeout 60 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "config5 and never" 2>&1 | tail -3 | tee $O/r2_final_dh88.log
echo "rc=$?"
