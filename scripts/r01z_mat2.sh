python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r01z11_gpu_tests.log 2>&1; tail -2 gpurun_out/r01z11_gpu_tests.log
B="python bench.py --skip-cpu --skip-train --skip-longform --steps 20"
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', round(d['ms_per_step'],3), d['gpu_launches_per_step'])"; }
for M in 0 1 0 1; do
TDANET_MAT2=$M $B 2>/dev/null | show "mat2 $M" >> gpurun_out/r01z11_mat2.txt
done
TDANET_MAT_ROWS=32 $B 2>/dev/null | show "mat2 1 rows32" >> gpurun_out/r01z11_mat2.txt
TDANET_MAT_ROWS=8 $B 2>/dev/null | show "mat2 1 rows8" >> gpurun_out/r01z11_mat2.txt
cat gpurun_out/r01z11_mat2.txt
