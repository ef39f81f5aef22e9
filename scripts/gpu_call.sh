#!/bin/bash
# One gpurun call of round 2: parity tests, bench, TF32 peak, ncu evidence.  Usage (from the repo root, on the box):
#   bash scripts/gpu_call.sh <tag> [tests] [bench] [peak] [ncu_gemm] [ncu_list] [ncu_train] [refarm] [det] [hunt_24000]
# Every artefact lands in gpurun_out/<tag>_*.
tag=$1; shift
out=gpurun_out
mkdir -p $out
for what in "$@"; do
  case $what in
    smoke)
      # a hung kernel must not eat the call: stop here when the tiny forward does not come back
      timeout -k 5 240 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1
      rc=$?; echo "smoke rc=$rc $(tail -2 $out/${tag}_smoke.log)"
      if [ $rc -ne 0 ]; then exit $rc; fi;;
    tests)
      timeout 1500 python -m pytest tests -m gpu -q -s > $out/${tag}_gpu_tests.log 2>&1
      echo "tests rc=$? $(tail -1 $out/${tag}_gpu_tests.log)";;
    tests_new)
      timeout 1500 python -m pytest tests/test_gpu_headline.py tests/test_gpu_round2.py -m gpu -q -s > $out/${tag}_gpu_tests_new.log 2>&1
      echo "tests_new rc=$? $(tail -1 $out/${tag}_gpu_tests_new.log)";;
    bench)
      timeout 900 python bench.py --steps 20 --warmup 5 --detail-out $out/${tag}_bench_detail.json > $out/${tag}_bench.json 2> $out/${tag}_bench.err
      echo "bench rc=$?"; cat $out/${tag}_bench.json;;
    bench_quick)
      timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-eager --skip-longform --skip-2ms --detail-out $out/${tag}_benchq_detail.json > $out/${tag}_benchq.json 2> $out/${tag}_benchq.err
      echo "bench_quick rc=$?"; cat $out/${tag}_benchq.json;;
    bench_bf16)
      timeout 600 python bench.py --steps 20 --warmup 5 --act-dtype bf16 --skip-cpu --skip-eager --skip-longform --skip-train --detail-out $out/${tag}_bf16_detail.json > $out/${tag}_bf16.json 2> $out/${tag}_bf16.err
      echo "bench_bf16 rc=$?"; cat $out/${tag}_bf16.json;;
    bench_fork)
      timeout 600 python bench.py --steps 20 --warmup 5 --variant fork --skip-cpu --skip-eager --skip-longform --detail-out $out/${tag}_fork_detail.json > $out/${tag}_fork.json 2> $out/${tag}_fork.err
      echo "bench_fork rc=$?"; cat $out/${tag}_fork.json;;
    refarm)
      timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_refarm.json 2> $out/${tag}_refarm.err
      echo "refarm rc=$?"; cat $out/${tag}_refarm.json;;
    peak)
      python scripts/measure_tf32_peak.py $out/${tag}_tf32_peak.json;;
    ncu_gemm)
      # one block's worth of tcgen05 GEMM launches (proj, in_proj, out_proj, fc1, fc2, res_conv), third block
      cmd="python bench.py --steps 1 --warmup 3 --no-graph --skip-train --skip-longform --skip-cpu --skip-eager --skip-2ms --detail-out $out/${tag}_ncu_dummy.json"
      $cmd > $out/${tag}_ncu_gemm_plain.log 2>&1 &&
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 12 -c 6 -f -o $out/${tag}_gemm $cmd > $out/${tag}_ncu_gemm.log 2>&1
      echo "ncu_gemm rc=$?";;
    ncuk_*)
      # ncu --set full of the launches whose name matches a regex (third block on): ncuk_<regex>[:skip[:count]]
      spec=${what#ncuk_}; rx=${spec%%:*}; rest=${spec#*:}; skip=12; cnt=6
      if [ "$rest" != "$spec" ]; then skip=${rest%%:*}; c2=${rest#*:}; if [ "$c2" != "$rest" ]; then cnt=$c2; fi; fi
      cmd="python bench.py --steps 1 --warmup 3 --no-graph --skip-train --skip-longform --skip-cpu --skip-eager --skip-2ms --detail-out $out/${tag}_ncu_dummy.json"
      timeout -k 5 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o $out/${tag}_k_${rx} $cmd > $out/${tag}_ncuk_${rx}.log 2>&1
      echo "ncuk $rx rc=$?";;
    ncu_list)
      cmd="python bench.py --steps 1 --warmup 3 --no-graph --skip-train --skip-longform --skip-cpu --skip-eager --skip-2ms --detail-out $out/${tag}_ncu_dummy.json"
      $cmd > $out/${tag}_ncu_list_plain.log 2>&1 &&
      timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1900 -c 480 --csv --log-file $out/${tag}_launches.csv $cmd > $out/${tag}_ncu_list.log 2>&1
      echo "ncu_list rc=$?";;
    ncu_train)
      cmd="python bench.py --train-only --steps 1 --warmup 3 --no-graph --skip-cpu"
      $cmd > $out/${tag}_ncu_train_plain.log 2>&1 &&
      timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 8000 -c 2000 --csv --log-file $out/${tag}_train_launches.csv $cmd > $out/${tag}_ncu_train.log 2>&1
      echo "ncu_train rc=$?";;
    ab_*)
      # A/B of an environment knob on the quick bench: ab_NAME=VALUE (e.g. ab_TDANET_LASTREAM_V=2)
      kv=${what#ab_}
      env $kv timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-eager --skip-longform --skip-2ms --detail-out $out/${tag}_${kv}_detail.json > $out/${tag}_${kv}.json 2> $out/${tag}_${kv}.err
      echo "ab $kv rc=$?"; python scripts/show_line.py $out/${tag}_${kv}.json;;
    abinf_*)
      kv=${what#abinf_}
      env $kv timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-eager --skip-longform --skip-2ms --skip-train --detail-out $out/${tag}_${kv}_detail.json > $out/${tag}_${kv}.json 2> $out/${tag}_${kv}.err
      echo "abinf $kv rc=$?"; python scripts/show_line.py $out/${tag}_${kv}.json;;
    train_bf16)
      timeout 600 python bench.py --train-only --steps 30 --warmup 5 --skip-cpu --act-dtype bf16 > $out/${tag}_train_bf16.json 2> $out/${tag}_train_bf16.err
      echo "train_bf16 rc=$?"; python -c "
import json
d=json.loads(open('$out/${tag}_train_bf16.json').read().strip().splitlines()[-1])
print('train bf16', d['value'], 'steps/s', d['ms_per_step'], 'ms; e2e', d['e2e']['value'], 'loss_after', d['loss_after'])
";;
    abfork_*)
      kv=${what#abfork_}
      fn=$(echo "$kv" | tr '/' '_')
      env $kv timeout 600 python bench.py --train-only --variant fork --steps 30 --warmup 5 --skip-cpu > $out/${tag}_forktrain_${fn}.json 2> $out/${tag}_forktrain_${fn}.err
      echo "abfork $kv rc=$?"; python -c "
import json
d=json.loads(open('$out/${tag}_forktrain_${fn}.json').read().strip().splitlines()[-1])
print('fork train', d['value'], 'steps/s', d['ms_per_step'], 'ms')
";;
    abtrain_*)
      kv=${what#abtrain_}
      fn=$(echo "$kv" | tr '/' '_')
      env $kv timeout 600 python bench.py --train-only --steps 30 --warmup 5 --skip-cpu > $out/${tag}_train_${fn}.json 2> $out/${tag}_train_${fn}.err
      echo "abtrain $kv rc=$?"; python -c "
import json,sys
d=json.loads(open('$out/${tag}_train_${fn}.json').read().strip().splitlines()[-1])
print('train', d['value'], 'steps/s', d['ms_per_step'], 'ms; e2e', d['e2e']['value'], 'launches', d['gpu_launches_per_step'], 'step_frac', d.get('step_frac'))
";;
    det)
      timeout 300 python scripts/det_probe.py $out/${tag}_deterministic.json > $out/${tag}_det.log 2>&1
      echo "det rc=$? $(tail -1 $out/${tag}_det.log | cut -c1-600)";;
    hunt_*)
      # rare run-to-run deviations with exact statistics: hunt_<runs> back-to-back 1-block forwards (scripts/det_hunt.py),
      # then the element-wise form on a third as many (scripts/det_hunt2.py); profiles/r04_bulk_ring_race.txt
      n=${what#hunt_}
      timeout 300 python scripts/det_hunt.py --blocks 1 --runs $n --out $out/${tag}_hunt_b1.json > $out/${tag}_hunt.log 2>&1
      echo "hunt rc=$? $(tail -1 $out/${tag}_hunt.log | cut -c1-600)"
      timeout 300 python scripts/det_hunt2.py --runs $((n / 3)) --out $out/${tag}_hunt2.json > $out/${tag}_hunt2.log 2>&1
      echo "hunt2 rc=$? $(tail -1 $out/${tag}_hunt2.log | cut -c1-600)";;
    *) echo "unknown step $what";;
  esac
done
