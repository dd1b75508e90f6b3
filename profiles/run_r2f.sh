#!/bin/bash
# round 2, 1-GPU call: what bounds the epilogue of conv_gemm_pair2?  Same launches with parts of the epilogue's memory traffic
# switched off (SACB_EPI2_DEBUG; results are wrong by construction, only the durations matter), then the L2 prefetch A/B.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_staged_epilogue_gpu.py -q > $O/r2f_pytest_epi.log 2>&1; echo "epi rc=$?"; tail -2 $O/r2f_pytest_epi.log
for dbg in 0 4; do
  SACB_EPI2_DEBUG=$dbg timeout 200 python profiles/conv_shapes.py epilogues > $O/r2f2_epilogues_dbg$dbg.txt 2>&1; echo "== SACB_EPI2_DEBUG=$dbg (bit2: no L2 prefetch of the next tile)"; grep -E "C256 K1024|C1024 K256" $O/r2f2_epilogues_dbg$dbg.txt
done
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
timeout 200 $B > $O/r2f2_bench.json 2> $O/r2f2_bench.err; echo "bench rc=$?"; cut -c1-200 $O/r2f2_bench.json
SACB_EPI2_DEBUG=4 timeout 200 $B > $O/r2f2_bench_nol2pf.json 2> $O/r2f2_bench_nol2pf.err; echo "bench (no L2 prefetch) rc=$?"; cut -c1-200 $O/r2f2_bench_nol2pf.json
SACB_EPI2=0 timeout 200 $B > $O/r2f2_bench_epi2off.json 2> $O/r2f2_bench_epi2off.err; echo "bench (EPI2 off) rc=$?"; cut -c1-200 $O/r2f2_bench_epi2off.json
