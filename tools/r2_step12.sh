#!/bin/bash
O=gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x -k "gemm or golden or config1 or attention or config2" 2>&1 | tail -8 | tee $O/r2_tests12.log
grep -q " failed\| error" $O/r2_tests12.log && exit 1
echo "== gemm microbench"; timeout 600 python tools/gemm_bench3.py quick 2>&1 | grep -E "oneshot|2cta" | tee $O/r2_gemm_bench3_duo.log
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-roofline > $O/r2_ab8.json 2> $O/r2_ab8.err
PSAM_GEMM_VARIANT=0x80 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-roofline > $O/r2_ab8_one.json 2> $O/r2_ab8_one.err
python - <<PY
import json
for n in ("r2_ab8","r2_ab8_one"):
    try:
        d=json.loads(open("$O/"+n+".json").read().strip().splitlines()[-1])
        print(n, round(d["value"],1), "clouds/s  e2e", round(d["e2e"]["value"],1), " c3", round(d["c3"]["value"],1), "clk", d["clocks"]["sm_mhz"])
    except Exception as e:
        print(n, "FAILED", e, open("$O/"+n+".err").read()[-500:])
PY
