#!/bin/bash
# round-1 final GPU call: parity suite with the L2 epilogue prefetch on, A/B bench, cuDNN baseline, ncu metrics of the streaming kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 120 python -m pytest tests -m gpu -x -q > $O/pytest_r1p.log 2>&1; echo "pytest rc=$?" >> $O/pytest_r1p.log; tail -3 $O/pytest_r1p.log
timeout 120 python bench.py > $O/bench_r1p_pf1.json 2> $O/bench_r1p_pf1.err; echo "pf1 rc=$?"; cat $O/bench_r1p_pf1.json | cut -c1-400
SACB_EPI_PREFETCH=0 timeout 120 python bench.py --no-cpu-baseline > $O/bench_r1p_pf0.json 2> $O/bench_r1p_pf0.err; echo "pf0 rc=$?"; cat $O/bench_r1p_pf0.json | cut -c1-200
timeout 150 python bench.py --impl reference --ref-device cuda --steps 4 --warmup 2 > $O/bench_cudnn_r1p.json 2> $O/bench_cudnn_r1p.err; echo "cudnn rc=$?"; cat $O/bench_cudnn_r1p.json | cut -c1-300; tail -3 $O/bench_cudnn_r1p.err
timeout 150 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none --profile-from-start off -k 'regex:tail_|loss_|sgd_|ema_|stem_im2col|maxpool|wgrad_finalize|prepare_batched|colsum|aspp_g|scatter2' \
  --csv --log-file $O/stream_kernels_r1p.csv python profiles/run_step.py 8 2 > $O/ncu_r1p.log 2>&1; echo "ncu rc=$?"; wc -l $O/stream_kernels_r1p.csv
SACB_EPI_PREFETCH=2 timeout 100 python bench.py --no-cpu-baseline > $O/bench_r1p_pf2.json 2> $O/bench_r1p_pf2.err; echo "pf2 rc=$?"; cat $O/bench_r1p_pf2.json | cut -c1-200
