show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', 'inf', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'train', d.get('train',{}).get('value'), 'longform', d.get('longform',{}).get('value'))"; }
python bench.py --skip-cpu 2>/dev/null | show default >> gpurun_out/r01z5_full.txt
TDANET_TILE_TARGET=592 TDANET_TILE_CAP=128 python bench.py --skip-cpu 2>/dev/null | show t592c128 >> gpurun_out/r01z5_full.txt
TDANET_TILE_TARGET=592 TDANET_TILE_CAP=128 python bench.py --skip-cpu --enc-ms 2 2>/dev/null | show 2ms_t592c128 >> gpurun_out/r01z5_full.txt
python bench.py --skip-cpu --enc-ms 2 2>/dev/null | show 2ms_default >> gpurun_out/r01z5_full.txt
TDANET_TILE_TARGET=592 TDANET_TILE_CAP=128 python bench.py --skip-cpu --variant fork 2>/dev/null | show fork_t592c128 >> gpurun_out/r01z5_full.txt
python bench.py --skip-cpu --variant fork 2>/dev/null | show fork_default >> gpurun_out/r01z5_full.txt
cat gpurun_out/r01z5_full.txt
