#!/bin/bash
# round 2, 2-GPU call (gpurun --gpus 2): correctness of everything that needs real ranks, then the N=2 bench lines.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
export SACB_P2P_TIMEOUT_S=30        # a wedged exchange must fail within the step's own timeout, not hang the box
nvidia-smi topo -m > $O/r2b_topo.txt 2>&1
SACB_RUN_UNVERIFIED=1 timeout 400 python -m pytest tests/test_p2p_gpu.py -q -s > $O/r2b_pytest_p2p.log 2>&1; echo "p2p rc=$?"; tail -6 $O/r2b_pytest_p2p.log
timeout 600 python -m pytest tests/test_world2_gpu.py -q -s > $O/r2b_pytest_world2.log 2>&1; echo "world2 rc=$?"; tail -12 $O/r2b_pytest_world2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2b_bench_n2.json 2> $O/r2b_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-250 $O/r2b_bench_n2.json; grep -i "exchange" $O/r2b_bench_n2.err | tail -3
SACB_NVLS=1 NCCL_DEBUG=WARN timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2b_bench_n2_nvls.json 2> $O/r2b_bench_n2_nvls.err; echo "bench n2 nvls rc=$?"; cut -c1-250 $O/r2b_bench_n2_nvls.json; grep -i "exchange\|error" $O/r2b_bench_n2_nvls.err | tail -5
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-p2p > $O/r2b_bench_n2_nccl.json 2> $O/r2b_bench_n2_nccl.err; echo "bench n2 nccl rc=$?"; cut -c1-250 $O/r2b_bench_n2_nccl.json
