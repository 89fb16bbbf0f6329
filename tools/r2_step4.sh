#!/bin/bash
O=gpurun_out
mkdir -p $O
echo "== fps tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "fps" 2>&1 | tail -15 | tee $O/r2_fps_tests.log
grep -q "failed" $O/r2_fps_tests.log && exit 1
echo "== tokenizer sweep"; timeout 300 python tools/tokenizer_sweep.py 2>&1 | tee $O/r2_tokenizer_sweep_b.md | tail -12
B="python bench.py --steps 30 --warmup 3 --no-c3 --no-cpu-baseline --no-gpu-reference --no-roofline"
for rep in 1 2 3; do
for cfg in "fold_dual::" "fold_nodual:PSAM_GEMM_VARIANT=0x10:" "nofold_dual:PSAM_FUSED_BLOCK_LN=0:" "nofold_nodual:PSAM_FUSED_BLOCK_LN=0 PSAM_GEMM_VARIANT=0x10:"; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}
  env $envs timeout 400 $B > $O/r2_ab3_${name}_$rep.json 2> $O/r2_ab3_${name}_$rep.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2_ab3_${name}_$rep.json").read().strip().splitlines()[-1])
    print("$rep $name", round(d["value"],1), "clouds/s  e2e", round(d["e2e"]["value"],1), " single-stream ms", round(d["run"]["single_stream_ms_per_cloud"],3), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name FAILED", e); print(open("$O/r2_ab3_${name}_$rep.err").read()[-800:])
PY
done
done
