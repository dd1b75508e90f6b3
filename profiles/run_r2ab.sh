#!/bin/bash
# round 2, 1-GPU call: ncu --set full (with source) of the streaming kernels of one step that sit furthest from their HBM floor:
# tail_probs / tail_pool / tail_refine, loss_fwd / loss_grad_rows, wgrad_finalize_batched.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off \
  -k regex:"tail_probs|tail_pool_kernel|tail_refine|loss_fwd|loss_grad_rows|wgrad_finalize_batched" -f -o $O/r2ab_stream \
  python profiles/run_step.py 8 2 > $O/r2ab_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $O/r2ab_ncu.log; ls -la $O/r2ab_stream.ncu-rep
