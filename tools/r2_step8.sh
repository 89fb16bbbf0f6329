#!/bin/bash
O=gpurun_out
mkdir -p $O
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x -k "knn or gemm or golden or config1 or group or interp" 2>&1 | tail -12 | tee $O/r2_tests8.log
grep -q " failed\| error" $O/r2_tests8.log && exit 1
echo "== pe gemm microbench"; timeout 300 python tools/gemm_bench3.py pe 2>&1 | tee $O/r2_gemm_bench3_pe.log | tail -14
echo "== tokenizer sweep"; timeout 300 python tools/tokenizer_sweep.py 2>&1 | tee $O/r2_tokenizer_sweep.md | tail -10
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-roofline"
for cfg in "default::" "twocta:PSAM_GEMM_VARIANT=0x1:" "dual:PSAM_GEMM_VARIANT=0x8:" "tiles1:PSAM_ATTENTION_TILES=1:" ; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}
  env $envs timeout 500 $B > $O/r2_ab6_${name}.json 2> $O/r2_ab6_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2_ab6_${name}.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"],1), "clouds/s  e2e", round(d["e2e"]["value"],1), " c3", round(d["c3"]["value"],1), "launches", d["launches_per_cloud"], "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name FAILED", e); print(open("$O/r2_ab6_${name}.err").read()[-800:])
PY
done
