#!/bin/bash
# round 2, last 1-GPU verification of the final tree: the whole GPU suite, smoke(), and the bench line exactly as the driver runs it.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r2af_pytest.log 2>&1; echo "suite rc=$?"; tail -3 $O/r2af_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2af_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2af_smoke.log
timeout 400 python bench.py > $O/r2af_bench.json 2> $O/r2af_bench.err; echo "bench rc=$?"; cut -c1-260 $O/r2af_bench.json
