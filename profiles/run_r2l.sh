#!/bin/bash
# round 2, 1-GPU call: conv_gemm_pair2 with the operand ring of four 32 KB units (was two 64 KB stages); roofline leg of bench.py
# with the streams serialised.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests/test_staged_epilogue_gpu.py tests/test_conv_gpu.py tests/test_step_gpu.py -q > $O/r2l_pytest.log 2>&1; echo "tests rc=$?"; tail -2 $O/r2l_pytest.log
timeout 200 python profiles/conv_shapes.py epilogues > $O/r2l_epilogues.txt 2>&1; cat $O/r2l_epilogues.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2l_bench.json 2> $O/r2l_bench.err; echo "bench rc=$?"; cut -c1-200 $O/r2l_bench.json; tail -2 $O/r2l_bench.err
