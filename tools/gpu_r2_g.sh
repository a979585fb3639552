#!/bin/bash
# round 2, call G: warp-specialised TMA Adam (parity + A/B), Level-1 integration test, BFS v5 profile
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_updates_gpu.py -x -q -m gpu > $O/g_updates.log 2>&1
echo "updates rc=$?" >> $O/g_updates.log
for ap in ldg ws16 ws8; do
  timeout 600 python bench.py --phase adam --adam-path $ap --steps 20 --warmup 3 > $O/g_adam_$ap.json 2> $O/g_adam_$ap.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bfs_kernel" -s 1 -c 1 -o $O/g_prof_bfs -f \
    python bench.py --phase bfs --bfs-roots 296 --steps 1 --warmup 1 > $O/g_ncu_bfs.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"adam_ws_kernel" -s 2 -c 1 -o $O/g_prof_adam_ws -f \
    python bench.py --phase adam --adam-path ws16 --steps 2 --warmup 2 > $O/g_ncu_adam.log 2>&1
tail -n 4 $O/g_updates.log
for ap in ldg ws16 ws8; do python -c "
import json
d=json.load(open('$O/g_adam_$ap.json')); print('$ap', d['ms_per_step'], d['roofline']['frac'])"; done
