#!/bin/bash
# round 2, 1-GPU call: first run of the residual-through-the-tensor-core route of conv_gemm_pair2 (RES_TENSOR) and of the BN
# scale folded into the fprop weight planes; two-stream modes now default.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests/test_staged_epilogue_gpu.py -q -s > $O/r2j_pytest_epi.log 2>&1; echo "epi rc=$?"; grep -E "case|passed|failed|Error|assert" $O/r2j_pytest_epi.log | tail -20
timeout 600 python -m pytest tests -m gpu -q -x > $O/r2j_pytest.log 2>&1; echo "suite rc=$?"; tail -5 $O/r2j_pytest.log
timeout 200 python profiles/conv_shapes.py epilogues > $O/r2j_epilogues.txt 2>&1; cat $O/r2j_epilogues.txt
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
timeout 200 $B > $O/r2j_bench.json 2> $O/r2j_bench.err; echo "bench rc=$?"; cut -c1-200 $O/r2j_bench.json; tail -2 $O/r2j_bench.err
SACB_EPI2_DEBUG=4 timeout 200 python profiles/conv_shapes.py epilogues > $O/r2j_epilogues_nopf.txt 2>&1; echo "== no L2 prefetch of residual tiles"; grep -E "res" $O/r2j_epilogues_nopf.txt
NCU="ncu --set full --import-source on --clock-control none -k regex:conv_gemm_pair -s 2 -c 1 -f"
timeout 200 $NCU -o $O/r2j_ncu_pair2_dgrad_res python profiles/conv_shapes.py one model.layer3.1.conv1 dgrad_res > $O/r2j_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 200 $NCU -o $O/r2j_ncu_pair2_fprop_res_unit python profiles/conv_shapes.py one model.layer3.1.conv3 fprop_res_unit > $O/r2j_ncu1.log 2>&1; echo "ncu1 rc=$?"
