B="python bench.py --skip-cpu --skip-train --skip-longform --steps 20"
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k={x['kernel']:x['ms_per_step'] for x in d['kernels']}
print('$1', round(d['ms_per_step'],3), 'spp_dw_s2', k.get('spp_dw_s2'), 'spp_dw0', k.get('spp_dw0'))"; }
for LB in 0 4; do for M in 0 8 16 32; do
TDANET_POOL_LB=$LB TDANET_POOL_MINROWS=$M $B 2>/dev/null | show "lb$LB min$M" >> gpurun_out/r01z_pool_sweep.txt
done; done
cat gpurun_out/r01z_pool_sweep.txt
