B="python bench.py --skip-cpu --skip-train --skip-longform --steps 20"
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', round(d['ms_per_step'],3))"; }
for M in 0 1 2 4 5 3 7 21 31; do
TDANET_SPP_REV=$M $B 2>/dev/null | show "spp_rev $M" >> gpurun_out/r01z10_l2.txt
done
cat gpurun_out/r01z10_l2.txt
