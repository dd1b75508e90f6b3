#!/bin/bash
# FIRST GPU call of round 2: verify what was written after round 1's GPU budget ran out, then A/B it.
#   gpurun --timeout 900 -- 'bash profiles/run_r2_first.sh'
# Everything lands in gpurun_out/ (r2a_*).  Each step has its own timeout; a trap in an unverified kernel only fails that step.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
# 0. the verified suite must still be green (default path; the unverified tests stay skipped here)
timeout 150 python -m pytest tests -m gpu -x -q > $O/r2a_pytest_default.log 2>&1; echo "default suite rc=$?"; tail -2 $O/r2a_pytest_default.log
# 1. unverified paths, one file at a time so that one failure does not hide the others
for t in test_abn_gpu test_staged_epilogue_gpu test_tail_split_gpu test_fast_mode_gpu; do
  SACB_RUN_UNVERIFIED=1 timeout 200 python -m pytest tests/$t.py -q -x > $O/r2a_$t.log 2>&1; echo "$t rc=$?"; tail -4 $O/r2a_$t.log
done
# 1b. two-stream teacher / student overlap (host-only change, SACB_TWO_STREAM=1): the full-step parity tests under the switch
SACB_TWO_STREAM=1 timeout 200 python -m pytest tests/test_step_gpu.py tests/test_variants_gpu.py -m gpu -q -x > $O/r2a_two_stream.log 2>&1; echo "two-stream suite rc=$?"; tail -2 $O/r2a_two_stream.log
# 2. bench A/B on one box: default, residual-staging epilogue, fast backward, fast everywhere
timeout 150 python bench.py --no-cpu-baseline > $O/r2a_bench_default.json 2> $O/r2a_bench_default.err; echo "bench default rc=$?"; cut -c1-200 $O/r2a_bench_default.json
SACB_EPI_STAGED=1 timeout 150 python bench.py --no-cpu-baseline > $O/r2a_bench_staged.json 2> $O/r2a_bench_staged.err; echo "bench staged rc=$?"; cut -c1-200 $O/r2a_bench_staged.json
SACB_TAIL_SPLIT=1 timeout 150 python bench.py --no-cpu-baseline > $O/r2a_bench_tail_split.json 2> $O/r2a_bench_tail_split.err; echo "bench tail-split rc=$?"; cut -c1-200 $O/r2a_bench_tail_split.json
SACB_TWO_STREAM=1 timeout 150 python bench.py --no-cpu-baseline > $O/r2a_bench_two_stream.json 2> $O/r2a_bench_two_stream.err; echo "bench two-stream rc=$?"; cut -c1-200 $O/r2a_bench_two_stream.json
SACB_PRECISION=fast_bwd timeout 150 python bench.py --no-cpu-baseline > $O/r2a_bench_fast_bwd.json 2> $O/r2a_bench_fast_bwd.err; echo "bench fast_bwd rc=$?"; cut -c1-200 $O/r2a_bench_fast_bwd.json
SACB_PRECISION=fast timeout 150 python bench.py --no-cpu-baseline > $O/r2a_bench_fast.json 2> $O/r2a_bench_fast.err; echo "bench fast rc=$?"; cut -c1-200 $O/r2a_bench_fast.json
