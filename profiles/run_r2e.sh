#!/bin/bash
# round 2, 1-GPU call: final form of conv_gemm_pair2 (8 warps, one prefetch set, LDS affine) re-verified; the two-stream
# variants (teacher || student forward; wgrad || dgrad in backward) tested and A/B-benched with 20 steps; the per-GPU shards of
# BASELINE.json configs[2], [3], [4]; the fast precision modes as separate lines.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_staged_epilogue_gpu.py -q > $O/r2e_pytest_epi.log 2>&1; echo "epi rc=$?"; tail -2 $O/r2e_pytest_epi.log
SACB_BWD_TWO_STREAM=1 timeout 400 python -m pytest tests/test_step_gpu.py tests/test_variants_gpu.py tests/test_abn_gpu.py -m gpu -q -x > $O/r2e_pytest_bwd2s.log 2>&1; echo "bwd-two-stream suite rc=$?"; tail -2 $O/r2e_pytest_bwd2s.log
timeout 200 python profiles/conv_shapes.py epilogues > $O/r2e_epilogues.txt 2>&1; cat $O/r2e_epilogues.txt
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
timeout 200 $B > $O/r2e_bench_default.json 2> $O/r2e_bench_default.err; echo "default rc=$?"; cut -c1-200 $O/r2e_bench_default.json
SACB_TWO_STREAM=1 timeout 200 $B > $O/r2e_bench_two_stream.json 2> $O/r2e_bench_two_stream.err; echo "two-stream rc=$?"; cut -c1-200 $O/r2e_bench_two_stream.json
SACB_BWD_TWO_STREAM=1 timeout 200 $B > $O/r2e_bench_bwd2s.json 2> $O/r2e_bench_bwd2s.err; echo "bwd-two-stream rc=$?"; cut -c1-200 $O/r2e_bench_bwd2s.json; tail -2 $O/r2e_bench_bwd2s.err
SACB_TWO_STREAM=1 SACB_BWD_TWO_STREAM=1 timeout 200 $B > $O/r2e_bench_both2s.json 2> $O/r2e_bench_both2s.err; echo "both rc=$?"; cut -c1-200 $O/r2e_bench_both2s.json
timeout 200 python bench.py --config 2 --groups 4 --steps 10 --warmup 3 > $O/r2e_cfg2_shard.json 2> $O/r2e_cfg2_shard.err; echo "cfg2 shard rc=$?"; cut -c1-220 $O/r2e_cfg2_shard.json; tail -2 $O/r2e_cfg2_shard.err
timeout 300 python bench.py --config 3 --groups 4 --steps 10 --warmup 3 > $O/r2e_cfg3_shard.json 2> $O/r2e_cfg3_shard.err; echo "cfg3 shard rc=$?"; cut -c1-220 $O/r2e_cfg3_shard.json; tail -2 $O/r2e_cfg3_shard.err
timeout 300 python bench.py --config 4 --groups 1 --steps 10 --warmup 3 > $O/r2e_cfg4_shard.json 2> $O/r2e_cfg4_shard.err; echo "cfg4 shard rc=$?"; cut -c1-220 $O/r2e_cfg4_shard.json; tail -2 $O/r2e_cfg4_shard.err
timeout 300 python bench.py --steps 12 --warmup 4 --no-cpu-baseline --also-fast > $O/r2e_bench_also_fast.json 2> $O/r2e_bench_also_fast.err; echo "also-fast rc=$?"; cut -c1-200 $O/r2e_bench_also_fast.json
