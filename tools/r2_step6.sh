#!/bin/bash
O=gpurun_out
mkdir -p $O
echo "== attention tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "fused_attention" 2>&1 | tail -25 | tee $O/r2_att3.log
grep -q "failed\|error" $O/r2_att3.log && exit 1
timeout 120 python tools/attention_trace.py 512 16 | tee $O/r2_att_trace2.log
bash tools/r2_step_ncu.sh attention
