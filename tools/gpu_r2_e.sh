#!/bin/bash
# round 2, call E: BFS v3 parity + timing, walk-kernel variants A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_walk_gpu.py tests/test_config_parity_gpu.py -x -q -m gpu > $O/e_walk.log 2>&1
echo "walk rc=$?" >> $O/e_walk.log
timeout 600 python bench.py --phase bfs --steps 5 --warmup 2 > $O/e_phase_bfs.json 2> $O/e_phase_bfs.err
timeout 2400 python tools/variants.py run > $O/e_variants.jsonl 2> $O/e_variants.err
tail -n 3 $O/e_walk.log
cut -c 1-500 $O/e_phase_bfs.json
cat $O/e_variants.jsonl
