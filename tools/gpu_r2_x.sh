#!/bin/bash
# round 2, call X: compute-sanitizer on the final kernels (flat steps on by default): memcheck + racecheck on the small fixtures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_walk_gpu.py -x -q -m gpu -k "tiny or rand300 or path_tail or hub_30k" > $O/x_memcheck.log 2>&1
echo "memcheck rc=$?" >> $O/x_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_walk_gpu.py -x -q -m gpu -k "walk and tiny" > $O/x_racecheck.log 2>&1
echo "racecheck rc=$?" >> $O/x_racecheck.log
tail -n 4 $O/x_memcheck.log $O/x_racecheck.log
grep -c "Race reported" $O/x_racecheck.log
grep "Race reported" $O/x_racecheck.log | sed 's/.*::\([a-z_0-9]*kernel\).* in \(.*\)$/\1 \2/' | sort | uniq -c | head
