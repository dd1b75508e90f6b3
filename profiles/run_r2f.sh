#!/bin/bash
# round 2, 1-GPU call: what bounds the epilogue of conv_gemm_pair2?  Same launches with parts of the epilogue's memory traffic
# switched off (SACB_EPI2_DEBUG; results are wrong by construction, only the durations matter).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
for dbg in 0 1 2 3; do
  SACB_EPI2_DEBUG=$dbg timeout 200 python profiles/conv_shapes.py epilogues > $O/r2f_epilogues_dbg$dbg.txt 2>&1; echo "== SACB_EPI2_DEBUG=$dbg (bit0: no residual/mask loads, bit1: no TMA stores)"; grep -E "C256 K1024|C1024 K256" $O/r2f_epilogues_dbg$dbg.txt
done
