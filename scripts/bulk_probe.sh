#!/bin/bash
# Which bulk-staged kernel breaks parity at which ring depth: the B = 64 headline parity test under a few settings.
out=gpurun_out; mkdir -p $out
for cfg in "TDANET_BULK_STAGES=2" "TDANET_BULK_STAGES=3 TDANET_BULK=3" "TDANET_BULK_STAGES=3 TDANET_BULK=5" "TDANET_BULK_STAGES=3 TDANET_BULK=9"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout -k 5 400 python -m pytest "tests/test_gpu_headline.py::test_headline_batch64_matches_oracle" -m gpu -q -x -s > $out/probe_$tag.log 2>&1
  echo "$cfg rc=$? $(grep -E 'passed|failed' $out/probe_$tag.log | tail -1) $(grep -o "AssertionError: {.*" $out/probe_$tag.log | head -1)"
done
