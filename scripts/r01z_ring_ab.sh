python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r01z3_gpu_tests.log 2>&1; tail -3 gpurun_out/r01z3_gpu_tests.log
B="python bench.py --skip-cpu --skip-train --skip-longform --steps 20"
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k={x['kernel']:x['ms_per_step'] for x in d['kernels']}
print('$1', round(d['ms_per_step'],3), 'spp_dw_s2', k.get('spp_dw_s2'), 'spp_dw0', k.get('spp_dw0'))"; }
for R in 0 1 0 1; do
TDANET_POOL_RING=$R $B 2>/dev/null | show "ring$R" >> gpurun_out/r01z3_ring.txt
done
TDANET_POOL_RING=1 TDANET_POOL_MINROWS=16 $B 2>/dev/null | show "ring1 min16" >> gpurun_out/r01z3_ring.txt
TDANET_POOL_RING=1 TDANET_TILE_TARGET=1184 $B 2>/dev/null | show "ring1 target1184" >> gpurun_out/r01z3_ring.txt
TDANET_POOL_RING=1 TDANET_TILE_TARGET=592 TDANET_TILE_CAP=128 $B 2>/dev/null | show "ring1 target592 cap128" >> gpurun_out/r01z3_ring.txt
cat gpurun_out/r01z3_ring.txt
