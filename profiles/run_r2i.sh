#!/bin/bash
# round 2, 4-GPU call: BASELINE.json configs[3] (VGG-16 FCN-8s, 16 groups x K=4 x 640^2 -> 4 groups per GPU) at its stated GPU count.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
export SACB_P2P_TIMEOUT_S=60
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521"
timeout 400 $TR bench.py --config 3 --gpus 4 --steps 20 --warmup 5 > $O/r2i_cfg3_n4.json 2> $O/r2i_cfg3_n4.err; echo "cfg3 n4 rc=$?"; cut -c1-250 $O/r2i_cfg3_n4.json; grep -i "exchange check" $O/r2i_cfg3_n4.err | tail -1
timeout 400 $TR bench.py --gpus 4 --steps 20 --warmup 5 > $O/r2i_cfg1_n4.json 2> $O/r2i_cfg1_n4.err; echo "cfg1 n4 rc=$?"; cut -c1-250 $O/r2i_cfg1_n4.json; grep -i "exchange check" $O/r2i_cfg1_n4.err | tail -1
