#!/bin/bash
# round 2, 8-GPU call: BASELINE.json configs[2] (32 groups x K=3 x 512^2 -> 4 groups per GPU) and configs[4] (8 groups x K=6 x
# 1024^2 -> 1 group per GPU) at their stated GPU count, fused peer-memory exchange with the bit-exactness check inside.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
export SACB_P2P_TIMEOUT_S=60
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
timeout 400 $TR bench.py --config 2 --gpus $N --steps 20 --warmup 5 > $O/r2h_cfg2_n$N.json 2> $O/r2h_cfg2_n$N.err; echo "cfg2 n$N rc=$?"; cut -c1-250 $O/r2h_cfg2_n$N.json; grep -i "exchange check" $O/r2h_cfg2_n$N.err | tail -1
timeout 400 $TR bench.py --config 4 --gpus $N --steps 20 --warmup 5 > $O/r2h_cfg4_n$N.json 2> $O/r2h_cfg4_n$N.err; echo "cfg4 n$N rc=$?"; cut -c1-250 $O/r2h_cfg4_n$N.json; grep -i "exchange check" $O/r2h_cfg4_n$N.err | tail -1
SACB_NVLS=1 timeout 400 $TR bench.py --config 2 --gpus $N --steps 20 --warmup 5 > $O/r2h_cfg2_n${N}_nvls.json 2> $O/r2h_cfg2_n${N}_nvls.err; echo "cfg2 n$N nvls rc=$?"; cut -c1-250 $O/r2h_cfg2_n${N}_nvls.json; grep -i "exchange check" $O/r2h_cfg2_n${N}_nvls.err | tail -1
