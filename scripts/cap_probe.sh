#!/bin/bash
# rows per CTA of the streaming kernels (one knob for all four): does a denser sweep of memory help?
out=gpurun_out; mkdir -p $out
for cap in 128 64 32 16; do
  env TDANET_LASTREAM_CAP=$cap TDANET_GSTATS_CAP=$cap TDANET_LSTATS_CAP=$cap TDANET_POOL_CAP=$cap timeout -k 5 400 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-eager --skip-longform --skip-2ms --skip-train --detail-out $out/cap_${cap}_detail.json > $out/cap_$cap.json 2> $out/cap_$cap.err
  echo "cap $cap rc=$?"; python scripts/show_line.py $out/cap_$cap.json
done
