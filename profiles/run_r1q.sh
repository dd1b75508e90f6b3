#!/bin/bash
# round-1 follow-up GPU call: parity suite + bench with the vectorised sgd / wgrad_finalize / tail_probs kernels,
# then the strict-fp32 PyTorch/cuDNN baseline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 100 python -m pytest tests -m gpu -x -q > $O/pytest_r1q.log 2>&1; echo "pytest rc=$?" >> $O/pytest_r1q.log; tail -3 $O/pytest_r1q.log
timeout 100 python bench.py > $O/bench_r1q.json 2> $O/bench_r1q.err; echo "bench rc=$?"; cut -c1-260 $O/bench_r1q.json
timeout 100 python bench.py --impl reference --ref-device cuda --steps 2 --warmup 1 --ref-strict-fp32 > $O/bench_cudnn_fp32_r1q.json 2> $O/bench_cudnn_fp32_r1q.err; echo "cudnn rc=$?"; cat $O/bench_cudnn_fp32_r1q.json | cut -c1-200; tail -2 $O/bench_cudnn_fp32_r1q.err
