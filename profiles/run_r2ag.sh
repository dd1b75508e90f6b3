#!/bin/bash
# round 2, 2-GPU call on the FINAL tree (after the layout-kernel and staged up-sampling changes): the two multi-GPU test
# modules and the N=2 bench line with the exchange check.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
export SACB_P2P_TIMEOUT_S=30
timeout 400 python -m pytest tests/test_p2p_gpu.py tests/test_world2_gpu.py -q -s > $O/r2ag_pytest_world2.log 2>&1; echo "world2 suites rc=$?"; grep -v "Warning\|symm_mem\|detach" $O/r2ag_pytest_world2.log | tail -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29527"
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2ag_bench_n2.json 2> $O/r2ag_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-250 $O/r2ag_bench_n2.json; grep -i "exchange check" $O/r2ag_bench_n2.err | tail -2
