#!/bin/bash
# round 2, final 1-GPU verification of the tree as committed: the whole GPU suite (with the printed numbers), smoke(), the bench
# line and the reference arm as the driver runs them, the ncu launch list of one step, the --set full captures the roofline
# fields quote, and the streaming-kernel DRAM table.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > $O/r2k_pytest.log 2>&1; echo "suite rc=$?"; tail -3 $O/r2k_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2k_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2k_smoke.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2k_bench.json 2> $O/r2k_bench.err; echo "bench rc=$?"; cut -c1-260 $O/r2k_bench.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2k_bench_ref.json 2> $O/r2k_bench_ref.err; echo "ref rc=$?"; cut -c1-200 $O/r2k_bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2k_launches_step.csv python profiles/run_step.py 8 2 > $O/r2k_launches.log 2>&1; echo "launch list rc=$?"
python profiles/launch_summary.py $O/r2k_launches_step.csv "round 2 final step (8 groups x K=3 x 512^2)" > $O/r2k_launches_summary.txt 2>&1; head -24 $O/r2k_launches_summary.txt
NCU="ncu --set full --import-source on --clock-control none -k regex:conv_gemm_pair -s 2 -c 1 -f"
timeout 200 $NCU -o $O/r2k_ncu_pair_3x3 python profiles/conv_shapes.py one model.layer3.1.conv2 fprop > $O/r2k_ncu0.log 2>&1; echo "ncu 3x3 rc=$?"
timeout 200 $NCU -o $O/r2k_ncu_pair2_fprop_res_unit python profiles/conv_shapes.py one model.layer3.1.conv3 fprop_res_unit > $O/r2k_ncu1.log 2>&1; echo "ncu 1x1+res rc=$?"
timeout 200 $NCU -o $O/r2k_ncu_pair2_dgrad_res python profiles/conv_shapes.py one model.layer3.1.conv1 dgrad_res > $O/r2k_ncu2.log 2>&1; echo "ncu dgrad_res rc=$?"
timeout 300 python profiles/conv_shapes.py > $O/r2k_conv_shapes.txt 2>&1; tail -3 $O/r2k_conv_shapes.txt
