#!/bin/bash
# round 2, call I: BFS v6 (512 threads x 16 entries, prefetched fixed slabs): sanitizer, parity, timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_walk_gpu.py -x -q -m gpu -k "bfs_matches and (tiny or rand300)" > $O/i_memcheck.log 2>&1
echo "memcheck rc=$?" >> $O/i_memcheck.log
timeout 900 python -m pytest tests/test_walk_gpu.py tests/test_config_parity_gpu.py -x -q -m gpu > $O/i_walk.log 2>&1
echo "walk rc=$?" >> $O/i_walk.log
timeout 600 python bench.py --phase bfs --steps 5 --warmup 2 > $O/i_phase_bfs.json 2> $O/i_phase_bfs.err
tail -n 3 $O/i_racecheck.log $O/i_memcheck.log $O/i_walk.log
cut -c 1-300 $O/i_phase_bfs.json
