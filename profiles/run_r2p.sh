#!/bin/bash
# round 2, last 1-GPU verification of the final tree: the whole GPU suite, smoke(), and the bench line exactly as the driver runs it.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > $O/r2p_pytest.log 2>&1; echo "suite rc=$?"; tail -3 $O/r2p_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2p_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2p_smoke.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2p_bench.json 2> $O/r2p_bench.err; echo "bench rc=$?"; cut -c1-260 $O/r2p_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2p_launches_step.csv python profiles/run_step.py 8 2 > $O/r2p_launches.log 2>&1; echo "launch list rc=$?"
python profiles/launch_summary.py $O/r2p_launches_step.csv "round 2 final tree, one step (8 groups x K=3 x 512^2)" > $O/r2p_launches_summary.txt 2>&1; head -14 $O/r2p_launches_summary.txt
