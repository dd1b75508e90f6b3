#!/bin/bash
# round 2, 8-GPU call on the final tree: the bench line of configs[1] (weak scaling, 8 groups per GPU) as the driver's scaling run does it.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
export SACB_P2P_TIMEOUT_S=60
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523"
timeout 400 $TR bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2r_cfg1_n8.json 2> $O/r2r_cfg1_n8.err; echo "cfg1 n8 rc=$?"; cut -c1-250 $O/r2r_cfg1_n8.json; grep -i "exchange check" $O/r2r_cfg1_n8.err | tail -1
