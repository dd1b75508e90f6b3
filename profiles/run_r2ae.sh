#!/bin/bash
# round 2, 1-GPU call: as run_r2ad.sh with loss_grad_rows2 templated on the outputs per thread and held to 64 registers (r2ad: 128 registers, 1.21 ms).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_tail_loss_variants_gpu.py tests/test_step_gpu.py -m gpu -q -p no:cacheprovider > $O/r2ae_pytest.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2ae_pytest.log
SACB_LIB=$PWD/da_sac_b200/libsac_b200_prev.so timeout 300 python profiles/ab_tail_loss.py > $O/r2ae_ab_prev.txt 2>&1; echo "probe prev rc=$?"
timeout 300 python profiles/ab_tail_loss.py > $O/r2ae_ab_new.txt 2>&1; echo "probe new rc=$?"
echo "digest lines that differ between the builds (losses are fp64 atomics, excluded):"
diff <(grep "^sha" $O/r2ae_ab_prev.txt | grep -v "loss.losses") <(grep "^sha" $O/r2ae_ab_new.txt | grep -v "loss.losses") | head -10; echo "diff rc=$?"
grep "^time" $O/r2ae_ab_prev.txt | sed 's/^/prev /'; grep "^time" $O/r2ae_ab_new.txt | sed 's/^/new  /'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2ae_launches.csv python profiles/run_step.py 8 2 > $O/r2ae_launches.log 2>&1; echo "launch list rc=$?"
python profiles/launch_summary.py $O/r2ae_launches.csv "staged up-sampling + segment loss backward, one step (8 groups x K=3 x 512^2)" > $O/r2ae_launches_summary.txt 2>&1
grep -E "total kernel|tail_probs|loss_fwd|loss_grad_rows|prepare_batched|finalize_batched" $O/r2ae_launches_summary.txt
