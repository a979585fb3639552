#!/bin/bash
# round 2, call L: level-synchronous walk steps (flat_*_kernel): sanitizer, parity with the steps on, A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
GG_FLAT_STEPS=3 timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_walk_gpu.py -x -q -m gpu -k "tiny or rand300" > $O/l_memcheck.log 2>&1
echo "memcheck rc=$?" >> $O/l_memcheck.log
GG_FLAT_STEPS=4 timeout 1200 python -m pytest tests/test_walk_gpu.py tests/test_config_parity_gpu.py -q -m gpu > $O/l_walk_flat4.log 2>&1
echo "walk rc=$?" >> $O/l_walk_flat4.log
for f in 0 2 4 6; do
  timeout 600 python bench.py --flat-steps $f --steps 20 --warmup 3 --no-cpu-baseline --g-steps 2 --verify 4 > $O/l_bench_flat$f.json 2> $O/l_bench_flat$f.err
done
tail -n 6 $O/l_memcheck.log $O/l_walk_flat4.log
for f in 0 2 4 6; do python - <<PY
import json
try:
    d=json.loads(open("$O/l_bench_flat$f.json").read().strip().splitlines()[-1])
    print("flat $f", round(d["value"]/1e6,2), "M/s e2e", round(d["e2e"]["value"]/1e6,2), d["parity"], d["roofline"]["k1_stage"]["walk_kernel_ms"], d["rates"]["g_mode"]["samples_per_s"])
except Exception as e:
    print("flat $f failed", e); print(open("$O/l_bench_flat$f.err").read()[-1500:])
PY
done
# ---- BFS top-down variants (A/B libraries built by hand: see DESIGN.md section 8)
for v in bfs_base bfs_amin bfs_defer bfs_both; do
  L=graphgan_b200/libgraphgan_b200.$v.so
  [ -f $L ] || continue
  GG_LIB=$L timeout 600 python -m pytest tests/test_walk_gpu.py -x -q -m gpu -k "bfs" > $O/l_bfs_$v.log 2>&1
  echo "$v pytest rc=$?" >> $O/l_bfs_$v.log
  GG_LIB=$L GG_BFS_BU_RATIO=0 timeout 600 python bench.py --phase bfs --steps 3 --warmup 1 > $O/l_phase_bfs_${v}_r0.json 2>> $O/l_bfs_$v.log
  GG_LIB=$L GG_BFS_BU_RATIO=1 timeout 600 python bench.py --phase bfs --steps 3 --warmup 1 > $O/l_phase_bfs_${v}_r1.json 2>> $O/l_bfs_$v.log
  tail -n 2 $O/l_bfs_$v.log
done
grep -o '"ms_per_root": [0-9.]*' $O/l_phase_bfs_*.json
