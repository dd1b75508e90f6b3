#!/bin/bash
# round 2, 1-GPU call: first run of conv_gemm_pair2_kernel (prefetch + TMA-store epilogue): bit-identity / fp64 test, the conv and
# step suites on top of it, per-shape timings, bench A/B against SACB_EPI2=0, one ncu --set full capture.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_staged_epilogue_gpu.py -q -s > $O/r2d_pytest_epi.log 2>&1; echo "epi rc=$?"; grep -v Warning $O/r2d_pytest_epi.log | tail -45
timeout 400 python -m pytest tests/test_conv_gpu.py tests/test_step_gpu.py tests/test_bench_geometry_gpu.py -q -x -s > $O/r2d_pytest_step.log 2>&1; echo "step rc=$?"; tail -4 $O/r2d_pytest_step.log
timeout 200 python profiles/conv_shapes.py epilogues > $O/r2d_epilogues.txt 2>&1; cat $O/r2d_epilogues.txt
SACB_EPI2=0 timeout 200 python profiles/conv_shapes.py epilogues > $O/r2d_epilogues_epi2off.txt 2>&1
timeout 200 python bench.py --steps 12 --warmup 4 --no-cpu-baseline > $O/r2d_bench.json 2> $O/r2d_bench.err; echo "bench rc=$?"; cut -c1-200 $O/r2d_bench.json; tail -2 $O/r2d_bench.err
SACB_EPI2=0 timeout 200 python bench.py --steps 12 --warmup 4 --no-cpu-baseline > $O/r2d_bench_epi2off.json 2> $O/r2d_bench_epi2off.err; echo "bench off rc=$?"; cut -c1-200 $O/r2d_bench_epi2off.json
NCU="ncu --set full --import-source on --clock-control none -k regex:conv_gemm_pair -s 2 -c 1 -f"
timeout 200 $NCU -o $O/r2d_ncu_pair2_1x1res python profiles/conv_shapes.py one model.layer3.1.conv3 fprop_res > $O/r2d_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 200 $NCU -o $O/r2d_ncu_pair2_dgrad_res python profiles/conv_shapes.py one model.layer3.1.conv1 dgrad_res > $O/r2d_ncu2.log 2>&1; echo "ncu2 rc=$?"
