set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r01z_gpu_tests.log 2>&1; tail -3 gpurun_out/r01z_gpu_tests.log
B="python bench.py --skip-cpu --skip-train --skip-longform --steps 20"
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k={x['kernel']:x['ms_per_step'] for x in d['kernels']}
print('$1', round(d['ms_per_step'],3), 'stats_global', k.get('la_stats_global'), 'first', k.get('la_combine_first'), 'comb', k.get('la_combine'))"; }
TDANET_GSTATS_STREAM=0 $B 2>/dev/null | show off >> gpurun_out/r01z_sweep.txt
$B 2>/dev/null | show default_592_128 >> gpurun_out/r01z_sweep.txt
for T in 296 444 888 1184; do for C in 64 128 256; do
TDANET_GSTATS_TARGET=$T TDANET_GSTATS_CAP=$C $B 2>/dev/null | show "t$T c$C" >> gpurun_out/r01z_sweep.txt
done; done
cat gpurun_out/r01z_sweep.txt
