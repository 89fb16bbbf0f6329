#!/bin/bash
O=gpurun_out
echo "== knn tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x -k "knn or golden or config4 or group or voronoi" 2>&1 | tail -8 | tee $O/r2_tests16.log
grep -q " failed\| error" $O/r2_tests16.log && exit 1
timeout 300 python tools/tokenizer_sweep.py 2>&1 | tail -9
timeout 300 ncu --metrics gpu__time_duration.sum -k regex:knn_kernel --clock-control none --profile-from-start off python tools/kernel_once.py knn_c4 2>&1 | grep -E "knn_kernel|duration" | tail -3
