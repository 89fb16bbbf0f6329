#!/bin/bash
# Validation of the attention slot-barrier fix (tight limits: the round's GPU budget is nearly spent).
O=gpurun_out
echo "== attention tests"; timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" 2>&1 | tail -3 | tee $O/r2_fc2_att.log
echo "== one query tile per CTA in the pipelined bench (hung before the fix)"
PSAM_ATTENTION_TILES=1 timeout 100 python bench.py --steps 6 --warmup 3 --no-c3 --no-cpu-baseline --no-gpu-reference --no-roofline > $O/r2_fc2_tiles1.json 2> $O/r2_fc2_tiles1.err; echo "rc=$?"; python -c "
import json; d=json.loads(open('$O/r2_fc2_tiles1.json').read().strip().splitlines()[-1]); print('tiles1', round(d['value'],1), round(d['e2e']['value'],1))" 2>&1 | tail -1
echo "== dh88 inside the model (opt-in path)"; PSAM_FUSED_ATTENTION_DH88=1 timeout 70 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "config5" 2>&1 | tail -3 | tee $O/r2_fc2_dh88.log
echo "== model tests"; timeout 150 python -m pytest tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -3 | tee $O/r2_fc2_model.log
