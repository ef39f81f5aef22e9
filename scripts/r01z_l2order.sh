python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r01z9_gpu_tests.log 2>&1; tail -2 gpurun_out/r01z9_gpu_tests.log
B="python bench.py --skip-cpu --skip-train --skip-longform --steps 20"
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', round(d['ms_per_step'],3))"; }
for M in 0 7 1 2 4 0 7; do
TDANET_L2_ORDER=$M $B 2>/dev/null | show "l2order $M" >> gpurun_out/r01z9_l2.txt
done
cat gpurun_out/r01z9_l2.txt
