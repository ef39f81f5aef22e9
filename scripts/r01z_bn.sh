B="python bench.py --skip-cpu --skip-train --skip-longform --steps 20"
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k={x['kernel']:x['ms_per_step'] for x in d['kernels']}
print('$1', round(d['ms_per_step'],3), 'proj', k.get('gemm_proj'), 'in_proj', k.get('gemm_in_proj'), 'fc1', k.get('gemm_fc1'), 'fc2', k.get('gemm_fc2'))"; }
for M in 256 128 256 128; do
TDANET_GEMM_BN=$M $B 2>/dev/null | show "bn $M" >> gpurun_out/r01z13_bn.txt
done
cat gpurun_out/r01z13_bn.txt
