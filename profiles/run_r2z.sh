#!/bin/bash
# round 2, 1-GPU call: (1) tests/test_ragged_gpu.py (non-square / odd-sized crops, K=1, everything ignored) on the B200;
# (2) A/B: how long a K loop conv_gemm_pair2_kernel (TMA-store epilogue, 2.5-stage unit ring) should take over from
#     conv_gemm_pair_kernel (3 stages, shuffle-transpose epilogue): SACB_EPI2_MAX_KB = 8 (default) / 16 (+ 1x1 1024->256) / 64 (all).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests/test_ragged_gpu.py -m gpu -q -s > $O/r2z_pytest_ragged.log 2>&1; echo "ragged rc=$?"; grep -E "passed|failed|logits rel|labels:" $O/r2z_pytest_ragged.log
for kb in 8 16 64; do
  SACB_EPI2_MAX_KB=$kb timeout 200 python profiles/conv_shapes.py epilogues > $O/r2z_epilogues_kb$kb.txt 2>&1; echo "== SACB_EPI2_MAX_KB=$kb"; cat $O/r2z_epilogues_kb$kb.txt | tail -12
done
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
for run in 1:8 2:16 3:64 4:8; do
  kb=${run#*:}; i=${run%%:*}
  SACB_EPI2_MAX_KB=$kb timeout 300 $B > $O/r2z_bench_${i}_kb$kb.json 2> $O/r2z_bench_${i}_kb$kb.err; echo "bench #$i kb=$kb rc=$?"; cut -c1-200 $O/r2z_bench_${i}_kb$kb.json
done
