#!/bin/bash
# GPU A/B bench pass (short runs), small logs only.
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 12 --warmup 3 --no-c3 --no-cpu-baseline --no-gpu-reference --no-roofline"
echo "== bench A/B"
for cfg in "new::" "noblockln:PSAM_FUSED_BLOCK_LN=0:" "nodual:PSAM_GEMM_VARIANT=0x10:" "twopass:PSAM_ATTENTION_TWOPASS=1:" "nodectc:PSAM_DECODER_TC=0:" "r1like:PSAM_FUSED_BLOCK_LN=0 PSAM_GEMM_VARIANT=0x10 PSAM_ATTENTION_TWOPASS=1 PSAM_DECODER_TC=0:"; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}
  env $envs timeout 400 $B > $O/r2_ab_$name.json 2> $O/r2_ab_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2_ab_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"],1), "clouds/s  e2e", round(d["e2e"]["value"],1), " single-stream ms", round(d["run"]["single_stream_ms_per_cloud"],3), "launches", d["launches_per_cloud"])
except Exception as e:
    print("$name FAILED", e); print(open("$O/r2_ab_$name.err").read()[-800:])
PY
done
echo "== depth sweep (new defaults)"
for d in 4 12; do timeout 300 $B --depth $d > $O/r2_depth$d.json 2> $O/r2_depth$d.err; python -c "
import json; d=json.loads(open('$O/r2_depth$d.json').read().strip().splitlines()[-1]); print('depth $d', round(d['value'],1))"; done
echo "== tokenizer sweep"; timeout 300 python tools/tokenizer_sweep.py 2>&1 | tee $O/r2_tokenizer_sweep.md | tail -14
echo "== full bench (default)"; timeout 1200 python bench.py > $O/r2_bench_c2.json 2> $O/r2_bench_c2.err; tail -c 600 $O/r2_bench_c2.err; head -c 6000 $O/r2_bench_c2.json
