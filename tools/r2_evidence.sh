#!/bin/bash
# Round-2 evidence session: full GPU test suite, ncu captures of the hot kernels (CSV pages only), launch list of one step,
# GEMM traffic, full bench line, side configs.  Everything lands in gpurun_out/ (small files).
O=gpurun_out
mkdir -p $O
echo "== tests"; timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee $O/r2_gpu_tests.log
grep -q " failed\| error" $O/r2_gpu_tests.log && exit 1
bash tools/r2_step_ncu.sh attention gemm_qkv gemm_m2048 gemm_pe gemm_rowln fps knn
rm -f $O/r2_gemm_qkv.source.csv $O/r2_gemm_m2048.source.csv $O/r2_gemm_pe.source.csv $O/r2_gemm_rowln.source.csv
bash tools/r2_step_launches.sh
python tools/ncu_traffic.py $O/r2_launches_c2.csv $O/r2_gemm_traffic.json
echo "== tokenizer sweep"; timeout 300 python tools/tokenizer_sweep.py 2>&1 | tee $O/r2_tokenizer_sweep.md | tail -12
echo "== full bench"; timeout 1500 python bench.py > $O/r2_bench_c2.json 2> $O/r2_bench_c2.err; tail -c 400 $O/r2_bench_c2.err; head -c 1500 $O/r2_bench_c2.json; echo
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_c2_reference_arm.json 2>/dev/null; head -c 600 $O/r2_bench_c2_reference_arm.json; echo
for c in c4 c5 c1; do
  timeout 600 python bench.py --config $c --depth 4 --steps 10 --warmup 3 --no-cpu-baseline > $O/r2_bench_$c.json 2> $O/r2_bench_$c.err
  python -c "
import json; d=json.loads(open('$O/r2_bench_$c.json').read().strip().splitlines()[-1]); print('$c', round(d['value'],1), 'clouds/s e2e', round(d['e2e']['value'],1), 'parity', d.get('gpu_reference',{}).get('parity'))" 2>&1 | tail -1
done
