python -m pytest tests -m gpu -x -q > gpurun_out/r01z8_gpu_tests.log 2>&1; tail -3 gpurun_out/r01z8_gpu_tests.log
B="python bench.py --skip-cpu --skip-train --skip-longform --steps 20"
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k={x['kernel']:x['ms_per_step'] for x in d['kernels']}
print('$1', round(d['ms_per_step'],3), 'attention', k.get('attention'), 'pool_sum', k.get('pool_sum'), 'misc', k.get('bottom_misc'))"; }
for i in 1 2; do
TDANET_ATT_RAW=0 TDANET_FUSE_LN_PE=0 $B 2>/dev/null | show "raw0 fuse0" >> gpurun_out/r01z8_ab.txt
TDANET_ATT_RAW=1 TDANET_FUSE_LN_PE=0 $B 2>/dev/null | show "raw1 fuse0" >> gpurun_out/r01z8_ab.txt
TDANET_ATT_RAW=0 TDANET_FUSE_LN_PE=1 $B 2>/dev/null | show "raw0 fuse1" >> gpurun_out/r01z8_ab.txt
$B 2>/dev/null | show "raw1 fuse1" >> gpurun_out/r01z8_ab.txt
done
cat gpurun_out/r01z8_ab.txt
