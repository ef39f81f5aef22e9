B="python bench.py --skip-cpu --skip-train --skip-longform --steps 20"
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k={x['kernel']:x['ms_per_step'] for x in d['kernels']}
print('$1', round(d['ms_per_step'],3), 'spp_dw_s2', k.get('spp_dw_s2'), 'spp_dw0', k.get('spp_dw0'), 'local', k.get('la_stats_local'), 'first', k.get('la_combine_first'))"; }
for T in 296 444 592 740 888; do for C in 64 128 256; do
TDANET_TILE_TARGET=$T TDANET_TILE_CAP=$C $B 2>/dev/null | show "target$T cap$C" >> gpurun_out/r01z4_tile.txt
done; done
cat gpurun_out/r01z4_tile.txt
