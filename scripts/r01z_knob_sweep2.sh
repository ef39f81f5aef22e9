B="python bench.py --skip-cpu --skip-train --skip-longform --steps 20"
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', round(d['ms_per_step'],3))"; }
$B 2>/dev/null | show "default" >> gpurun_out/r01z7_knobs.txt
for P in "592 128" "888 128" "592 256" "296 256" "1776 128"; do set -- $P
TDANET_LASTREAM_TARGET=$1 TDANET_LASTREAM_CAP=$2 $B 2>/dev/null | show "lastream $1/$2" >> gpurun_out/r01z7_knobs.txt
done
for P in "1184 64" "592 64" "592 128"; do set -- $P
TDANET_LAT_TARGET=$1 TDANET_LAT_CAP=$2 $B 2>/dev/null | show "lat $1/$2" >> gpurun_out/r01z7_knobs.txt
done
for M in 32 64; do
TDANET_MAT_ROWS=$M $B 2>/dev/null | show "mat $M" >> gpurun_out/r01z7_knobs.txt
done
cat gpurun_out/r01z7_knobs.txt
