#!/bin/bash
# round 2, call Q: split launches per flat step (hub items | flat_score_kernel) A/B, cleanup check
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
for sp in 0 1; do
  GG_FLAT_SPLIT=$sp timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --g-steps 2 --verify 4 > $O/q_bench_split$sp.json 2> $O/q_bench_split$sp.err
  GG_FLAT_SPLIT=$sp timeout 900 python -m pytest tests/test_walk_gpu.py tests/test_config_parity_gpu.py -q -m gpu -x > $O/q_pytest_split$sp.log 2>&1
  tail -n 1 $O/q_pytest_split$sp.log
  python - <<PY
import json
try:
    d=json.loads(open("$O/q_bench_split$sp.json").read().strip().splitlines()[-1])
    k=d["roofline"]["k1_stage"]
    print("split $sp", round(d["value"]/1e6,2), "M/s e2e", round(d["e2e"]["value"]/1e6,2), d["parity"]["mismatches"], d["ms_per_step"], k["hub_scores_root_cdf_ms"], k["root_step_step1_cdf_ms"], k["walk_kernel_ms"], k["finalize_emit_ms"], d["rates"]["g_mode"]["samples_per_s"])
except Exception as e:
    print("split $sp failed", e); print(open("$O/q_bench_split$sp.err").read()[-1500:])
PY
done
timeout 600 python bench.py --phase bfs --steps 5 --warmup 2 > $O/q_phase_bfs.json 2> $O/q_phase_bfs.err
grep -o '"ms_per_root": [0-9.]*' $O/q_phase_bfs.json
