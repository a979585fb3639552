#!/bin/bash
# round 2, call A: new BFS (tree bits) + bitmap walk: sanitizer on small cases, full GPU tests, quick bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/a_gpu.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_walk_gpu.py -x -q -m gpu -k "bfs_matches and (tiny or rand300)" > $O/a_memcheck.log 2>&1
echo "memcheck rc=$?" >> $O/a_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_walk_gpu.py -x -q -m gpu -k "bfs_matches and tiny" > $O/a_racecheck.log 2>&1
echo "racecheck rc=$?" >> $O/a_racecheck.log
timeout 1200 python -m pytest tests/test_walk_gpu.py -x -q -m gpu > $O/a_walk.log 2>&1
echo "walk rc=$?" >> $O/a_walk.log
timeout 1500 python -m pytest tests/test_config_parity_gpu.py -x -q -m gpu > $O/a_config.log 2>&1
echo "config rc=$?" >> $O/a_config.log
timeout 900 python -m pytest tests/test_updates_gpu.py tests/test_dist.py -x -q -m gpu > $O/a_updates.log 2>&1
echo "updates rc=$?" >> $O/a_updates.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/a_bench.json 2> $O/a_bench.err
echo "bench rc=$?" >> $O/a_bench.err
tail -3 $O/a_memcheck.log $O/a_racecheck.log $O/a_walk.log $O/a_config.log $O/a_updates.log
cat $O/a_bench.json | head -c 3000
