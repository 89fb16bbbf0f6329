#!/bin/bash
# Round-2 GPU session 1: correctness of everything new, then A/B benches and ncu captures.  Everything lands in gpurun_out/.
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 12 --warmup 3 --no-c3 --no-cpu-baseline --no-gpu-reference --no-roofline"
echo "== attention tests"; timeout 500 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "fused_attention" 2>&1 | tail -15 | tee $O/r2_att.log
echo "== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --deselect tests/test_gpu_kernels.py::test_fused_attention_tc 2>&1 | tail -25 | tee $O/r2_kernels.log
echo "== model tests"; timeout 1500 python -m pytest tests/test_gpu_model.py -m gpu -q -s 2>&1 | grep -E "parity|passed|failed|Error|error|assert" | tail -60 | tee $O/r2_model.log
echo "== bench A/B"
for cfg in "new::" "noblockln:PSAM_FUSED_BLOCK_LN=0:" "nodual:PSAM_GEMM_VARIANT=0x10:" "nodual_noblockln:PSAM_FUSED_BLOCK_LN=0 PSAM_GEMM_VARIANT=0x10:" "twopass:PSAM_ATTENTION_TWOPASS=1:" "nodectc:PSAM_DECODER_TC=0:" "r1like:PSAM_FUSED_BLOCK_LN=0 PSAM_GEMM_VARIANT=0x10 PSAM_ATTENTION_TWOPASS=1 PSAM_DECODER_TC=0:"; do
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
echo "== tokenizer sweep"; timeout 300 python tools/tokenizer_sweep.py 2>&1 | tee $O/r2_tokenizer_sweep.md | tail -12
echo "== full bench (default)"; timeout 900 python bench.py > $O/r2_bench_c2.json 2> $O/r2_bench_c2.err; tail -c 600 $O/r2_bench_c2.err; head -c 3000 $O/r2_bench_c2.json
echo "== ncu"
for k in attention gemm_qkv gemm_qkv_single fps knn; do
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -o $O/r2_$k -f python tools/kernel_once.py $k > $O/r2_ncu_$k.log 2>&1; tail -2 $O/r2_ncu_$k.log
done
timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_active.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file $O/r2_launches_c2.csv python tools/profile_step.py --config c2 --throughput-tiles > $O/r2_profile_step.log 2>&1; tail -2 $O/r2_profile_step.log
ls -la $O | tail -30
