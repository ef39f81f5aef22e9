for T in 592 1184 2368 4736 9472; do for C in 64 128; do
  TDANET_TILE_TARGET=$T TDANET_TILE_CAP=$C python bench.py --skip-cpu --skip-train --skip-longform --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k={x['kernel']:x['ms_per_step'] for x in d['kernels']}
print('target $T cap $C', round(d['ms_per_step'],3), 'spp_dw_s2', k.get('spp_dw_s2'), 'spp_dw0', k.get('spp_dw0'), 'stats_global', k.get('la_stats_global'))"
done; done
