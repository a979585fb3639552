#!/bin/bash
# round 2, call K: direction-optimising BFS (bottom_up_level): sanitizer, parity in every mode, ratio sweep, ncu
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_walk_gpu.py -x -q -m gpu -k "bottom_up_levels and (path_tail or hub_30k)" > $O/k_memcheck.log 2>&1
echo "memcheck rc=$?" >> $O/k_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_walk_gpu.py -x -q -m gpu -k "bottom_up_levels and path_tail" > $O/k_racecheck.log 2>&1
echo "racecheck rc=$?" >> $O/k_racecheck.log
timeout 900 python -m pytest tests/test_walk_gpu.py tests/test_config_parity_gpu.py -x -q -m gpu > $O/k_walk.log 2>&1
echo "walk rc=$?" >> $O/k_walk.log
for r in 0 1 1.5 2.5 4; do
  GG_BFS_BU_RATIO=$r timeout 600 python bench.py --phase bfs --steps 3 --warmup 1 > $O/k_phase_bfs_r$r.json 2> $O/k_phase_bfs_r$r.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bfs_kernel" -s 1 -c 1 -o $O/k_prof_bfs -f \
    python bench.py --phase bfs --bfs-roots 296 --steps 1 --warmup 1 > $O/k_ncu_bfs.log 2>&1
tail -n 4 $O/k_memcheck.log $O/k_racecheck.log $O/k_walk.log
grep -o '"ms_per_root": [0-9.]*' $O/k_phase_bfs_r*.json
