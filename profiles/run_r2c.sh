#!/bin/bash
# round 2, 1-GPU call: full GPU suite with the un-gated files and the new bench-geometry parity test, the bench line in its
# new form (teacher-update graph, self_ce, reference-arm cpu_baseline), the reference arm itself, and ncu --set full captures
# (with source) of the epilogue-bound launches of the pair kernel.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x > $O/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r2c_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2c_bench.json 2> $O/r2c_bench.err; echo "bench rc=$?"; cut -c1-300 $O/r2c_bench.json; tail -3 $O/r2c_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2c_bench_ref.json 2> $O/r2c_bench_ref.err; echo "ref rc=$?"; cut -c1-300 $O/r2c_bench_ref.json
NCU="ncu --set full --import-source on --clock-control none -k regex:conv_gemm_pair -s 2 -c 1 -f"
timeout 200 $NCU -o $O/r2c_ncu_pair1x1res python profiles/conv_shapes.py one model.layer3.1.conv3 fprop_res > $O/r2c_ncu1.log 2>&1; echo "ncu1 rc=$?"
SACB_EPI_STAGED=1 timeout 200 $NCU -o $O/r2c_ncu_pair1x1res_staged python profiles/conv_shapes.py one model.layer3.1.conv3 fprop_res > $O/r2c_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 200 $NCU -o $O/r2c_ncu_pair1x1_plain python profiles/conv_shapes.py one model.layer3.1.conv3 fprop > $O/r2c_ncu3.log 2>&1; echo "ncu3 rc=$?"
timeout 200 $NCU -o $O/r2c_ncu_pair_dgrad_res python profiles/conv_shapes.py one model.layer3.1.conv1 dgrad_res > $O/r2c_ncu4.log 2>&1; echo "ncu4 rc=$?"
timeout 200 $NCU -o $O/r2c_ncu_pair_1024to256 python profiles/conv_shapes.py one model.layer3.1.conv1 fprop > $O/r2c_ncu5.log 2>&1; echo "ncu5 rc=$?"
python profiles/conv_shapes.py epilogues > $O/r2c_epilogues.txt 2>&1; cat $O/r2c_epilogues.txt
ls -la $O/*.ncu-rep
