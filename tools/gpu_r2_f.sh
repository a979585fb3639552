#!/bin/bash
# round 2, call F: BFS v5, TMA hub staging, single-tile choose fast path: full GPU tests; adam A/B; bench lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests/test_walk_gpu.py tests/test_config_parity_gpu.py tests/test_updates_gpu.py -x -q -m gpu > $O/f_tests.log 2>&1
echo "tests rc=$?" >> $O/f_tests.log
timeout 600 python bench.py --phase bfs --steps 5 --warmup 2 > $O/f_phase_bfs.json 2> $O/f_phase_bfs.err
for ap in ldg tma tma256x2 tma512x3; do
  timeout 600 python bench.py --phase adam --adam-path $ap --steps 20 --warmup 3 > $O/f_adam_$ap.json 2> $O/f_adam_$ap.err
done
timeout 900 python bench.py --no-cpu-baseline --verify 6 > $O/f_bench.json 2> $O/f_bench.err
timeout 900 python bench.py --no-cpu-baseline --verify 0 --g-steps 0 --no-tma > $O/f_bench_notma.json 2> $O/f_bench_notma.err
tail -n 3 $O/f_tests.log
cut -c 1-260 $O/f_phase_bfs.json
for ap in ldg tma tma256x2 tma512x3; do python -c "
import json,sys
d=json.load(open('$O/f_adam_$ap.json')); print('$ap', d['ms_per_step'], d['roofline']['frac'])"; done
python -c "
import json
for f in ('f_bench','f_bench_notma'):
    d=json.load(open('$O/'+f+'.json')); print(f, d['value'], d['e2e']['value'], d['roofline']['k1_stage'], d['full_pass']['bfs_ms_per_root'], d['parity'])"
