#!/bin/bash
O=gpurun_out
mkdir -p $O
echo "== gemm variant tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm" 2>&1 | tail -30 | tee $O/r2_gemm_tests.log
echo "== model golden"; timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "golden or config1" 2>&1 | tail -30 | tee $O/r2_model_quick.log
echo "== gemm microbench"; timeout 900 python tools/gemm_bench3.py 2>&1 | tee $O/r2_gemm_bench3.log | tail -80
B="python bench.py --steps 12 --warmup 3 --no-c3 --no-cpu-baseline --no-gpu-reference --no-roofline"
for cfg in "new::" "noblockln:PSAM_FUSED_BLOCK_LN=0:" "noblockln_nodual:PSAM_FUSED_BLOCK_LN=0 PSAM_GEMM_VARIANT=0x10:" "noblockln_nopersist:PSAM_FUSED_BLOCK_LN=0 PSAM_GEMM_VARIANT=0x40:"; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}
  env $envs timeout 400 $B > $O/r2_ab2_$name.json 2> $O/r2_ab2_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2_ab2_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"],1), "clouds/s  e2e", round(d["e2e"]["value"],1), " single-stream ms", round(d["run"]["single_stream_ms_per_cloud"],3), "launches", d["launches_per_cloud"])
except Exception as e:
    print("$name FAILED", e); print(open("$O/r2_ab2_$name.err").read()[-800:])
PY
done
