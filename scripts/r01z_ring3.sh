python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r01z12_gpu_tests.log 2>&1; tail -2 gpurun_out/r01z12_gpu_tests.log
B="python bench.py --skip-cpu --skip-train --skip-longform --steps 20"
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k={x['kernel']:x['ms_per_step'] for x in d['kernels']}
print('$1', round(d['ms_per_step'],3), 'spp_dw_s2', k.get('spp_dw_s2'), 'spp_dw0', k.get('spp_dw0'))"; }
for M in 2 3 2 3; do
TDANET_POOL_RING=$M $B 2>/dev/null | show "ring $M" >> gpurun_out/r01z12_ring3.txt
done
TDANET_POOL_RING=3 TDANET_POOL_TARGET=444 $B 2>/dev/null | show "ring 3 target 444" >> gpurun_out/r01z12_ring3.txt
TDANET_POOL_RING=3 TDANET_POOL_TARGET=296 TDANET_POOL_CAP=256 $B 2>/dev/null | show "ring 3 target 296/256" >> gpurun_out/r01z12_ring3.txt
cat gpurun_out/r01z12_ring3.txt
