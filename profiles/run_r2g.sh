#!/bin/bash
# round 2, 2-GPU call: world-2 suites after the fixes (exchange check gathers the sharded momentum; running_conf bar), the N=2
# bench lines with the exchange check inside (plain peer-memory kernel, NVLS, two-stream variants).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
export SACB_P2P_TIMEOUT_S=30
timeout 400 python -m pytest tests/test_p2p_gpu.py tests/test_world2_gpu.py -q -s > $O/r2g_pytest_world2.log 2>&1; echo "world2 suites rc=$?"; grep -v "Warning\|symm_mem" $O/r2g_pytest_world2.log | tail -12
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2g_bench_n2.json 2> $O/r2g_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-250 $O/r2g_bench_n2.json; grep -i "exchange check" $O/r2g_bench_n2.err | tail -2
SACB_NVLS=1 timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2g_bench_n2_nvls.json 2> $O/r2g_bench_n2_nvls.err; echo "bench n2 nvls rc=$?"; cut -c1-250 $O/r2g_bench_n2_nvls.json; grep -i "exchange check\|Error" $O/r2g_bench_n2_nvls.err | tail -3
SACB_TWO_STREAM=1 SACB_BWD_TWO_STREAM=1 timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2g_bench_n2_both2s.json 2> $O/r2g_bench_n2_both2s.err; echo "bench n2 two-stream rc=$?"; cut -c1-250 $O/r2g_bench_n2_both2s.json; grep -i "exchange check\|Error" $O/r2g_bench_n2_both2s.err | tail -3
