#!/bin/bash
# round 2, 1-GPU call: tail split of the pair kernel again, now with ONE epilogue instantiation shared by full and half tiles.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_tail_split_gpu.py -q > $O/r2m_pytest.log 2>&1; echo "tail-split test rc=$?"; tail -2 $O/r2m_pytest.log
for ts in 0 1; do
  SACB_TAIL_SPLIT=$ts timeout 200 python profiles/conv_shapes.py epilogues > $O/r2m_epilogues_ts$ts.txt 2>&1; echo "== SACB_TAIL_SPLIT=$ts"; grep -E "3x3" $O/r2m_epilogues_ts$ts.txt
done
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
timeout 300 $B > $O/r2m_bench_ts0.json 2> $O/r2m_bench_ts0.err; echo "bench rc=$?"; cut -c1-200 $O/r2m_bench_ts0.json
SACB_TAIL_SPLIT=1 timeout 300 $B > $O/r2m_bench_ts1.json 2> $O/r2m_bench_ts1.err; echo "bench (tail split) rc=$?"; cut -c1-200 $O/r2m_bench_ts1.json
