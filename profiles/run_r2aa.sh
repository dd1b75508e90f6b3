#!/bin/bash
# round 2, 1-GPU call: tiled weight re-layout (sacb_prepare_batched) and staged-transpose filter-gradient finalize
# (sacb_wgrad_finalize_batched): direct bit-exactness tests + golden step tests, per-kernel times from an ncu launch list,
# same-box A/B of the step against the previous build (da_sac_b200/libsac_b200_prev.so = HEAD before this change).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_prepare_finalize_gpu.py tests/test_step_gpu.py -m gpu -q -p no:cacheprovider > $O/r2aa_pytest.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2aa_pytest.log
for v in prev new; do
  if [ $v = prev ]; then export SACB_LIB=$PWD/da_sac_b200/libsac_b200_prev.so; else unset SACB_LIB; fi
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2aa_launches_$v.csv python profiles/run_step.py 8 2 > $O/r2aa_launches_$v.log 2>&1; echo "launch list ($v) rc=$?"
  python profiles/launch_summary.py $O/r2aa_launches_$v.csv "$v build, one step (8 groups x K=3 x 512^2)" > $O/r2aa_launches_${v}_summary.txt 2>&1
  grep -E "total kernel|prepare_batched|finalize_batched" $O/r2aa_launches_${v}_summary.txt
done
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
for run in 1:prev 2:new 3:prev 4:new; do
  v=${run#*:}; i=${run%%:*}
  if [ $v = prev ]; then export SACB_LIB=$PWD/da_sac_b200/libsac_b200_prev.so; else unset SACB_LIB; fi
  timeout 300 $B > $O/r2aa_bench_${i}_$v.json 2> $O/r2aa_bench_${i}_$v.err; echo "bench #$i $v rc=$?"; cut -c1-200 $O/r2aa_bench_${i}_$v.json
done
