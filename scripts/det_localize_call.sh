# residual run-to-run differences with exact statistics: how often, from how many blocks on, under which knobs
o=gpurun_out
run() { tag=$1; shift; env "$@" timeout 150 python scripts/det_localize.py --out $o/r04d_loc_$tag.json $ARGS > $o/r04d_loc_$tag.log 2>&1; echo $tag rc=$? $(tail -1 $o/r04d_loc_$tag.log | cut -c1-900); }
ARGS="--blocks 2 --runs 300" run b2 X=1
ARGS="--blocks 16 --runs 60" run b16 X=1
ARGS="--blocks 16 --runs 60" run b16_nopdl TDANET_PDL=0
ARGS="--blocks 16 --runs 60" run b16_notma TDANET_GEMM_TMA_STORE=0
ARGS="--blocks 16 --runs 60" run b16_nobulk TDANET_BULK=0
