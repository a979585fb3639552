#!/bin/bash
# round 2, call S (2 GPUs): which change breaks save -> load -> continue under sharding?
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m pytest tests/test_dist.py -q -m gpu > $O/s_dist_$name.log 2>&1
  echo "$name rc=$?" | tee -a $O/s_dist_$name.log
  grep -E "AssertionError|passed|failed" $O/s_dist_$name.log | head -5
}
run default A=1
run noflat GG_FLAT_STEPS=0
run nobu GG_BFS_BU_RATIO=0
run nobu_noflat GG_BFS_BU_RATIO=0 GG_FLAT_STEPS=0
(cd _old && timeout 600 python -m pytest tests/test_dist.py -q -m gpu > ../$O/s_dist_old.log 2>&1; echo "old rc=$?" | tee -a ../$O/s_dist_old.log; grep -E "AssertionError|passed|failed" ../$O/s_dist_old.log | head -5)
