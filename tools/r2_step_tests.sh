#!/bin/bash
# GPU correctness pass: every -m gpu test, logs into gpurun_out/ (small files only).
O=gpurun_out
mkdir -p $O
echo "== attention tests"; timeout 500 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "fused_attention" 2>&1 | tail -40 | tee $O/r2_att.log
echo "== kernel tests"; timeout 1200 python -m pytest tests/test_gpu_kernels.py -m gpu -q --deselect tests/test_gpu_kernels.py::test_fused_attention_tc 2>&1 | tail -80 | tee $O/r2_kernels.log
echo "== model tests"; timeout 1800 python -m pytest tests/test_gpu_model.py -m gpu -q 2>&1 | tail -80 | tee $O/r2_model.log
