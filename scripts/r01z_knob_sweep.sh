B="python bench.py --skip-cpu --skip-train --skip-longform --steps 20"
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', round(d['ms_per_step'],3))"; }
for P in "2368 64" "592 128" "1184 128"; do for L in "2368 64" "592 128" "1184 128" "296 128"; do
set -- $P; PT=$1; PC=$2; set -- $L; LT=$1; LC=$2
TDANET_POOL_TARGET=$PT TDANET_POOL_CAP=$PC TDANET_LSTATS_TARGET=$LT TDANET_LSTATS_CAP=$LC $B 2>/dev/null | show "pool $PT/$PC lstats $LT/$LC" >> gpurun_out/r01z6_knobs.txt
done; done
cat gpurun_out/r01z6_knobs.txt
