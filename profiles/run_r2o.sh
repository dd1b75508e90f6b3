#!/bin/bash
# round 2, 1-GPU call: pair wgrad with one K range per CTA pair (half the split-K partials) A/B; 1-GPU shard of configs[2] with
# the final kernels (denominator of the 8-GPU efficiency).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests/test_conv_gpu.py tests/test_step_gpu.py tests/test_fast_mode_gpu.py -q > $O/r2o_pytest.log 2>&1; echo "tests rc=$?"; tail -2 $O/r2o_pytest.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
timeout 300 $B > $O/r2o_bench.json 2> $O/r2o_bench.err; echo "bench rc=$?"; cut -c1-200 $O/r2o_bench.json
SACB_WGRAD_ONE_WAVE=0 timeout 300 $B > $O/r2o_bench_two_waves.json 2> $O/r2o_bench_two_waves.err; echo "bench (two K ranges per pair) rc=$?"; cut -c1-200 $O/r2o_bench_two_waves.json
timeout 300 python bench.py --config 2 --groups 4 --steps 20 --warmup 5 > $O/r2o_cfg2_shard.json 2> $O/r2o_cfg2_shard.err; echo "cfg2 shard rc=$?"; cut -c1-220 $O/r2o_cfg2_shard.json
timeout 300 python bench.py --config 4 --groups 1 --steps 20 --warmup 5 > $O/r2o_cfg4_shard.json 2> $O/r2o_cfg4_shard.err; echo "cfg4 shard rc=$?"; cut -c1-220 $O/r2o_cfg4_shard.json
timeout 300 python bench.py --config 3 --groups 4 --steps 20 --warmup 5 > $O/r2o_cfg3_shard.json 2> $O/r2o_cfg3_shard.err; echo "cfg3 shard rc=$?"; cut -c1-220 $O/r2o_cfg3_shard.json
