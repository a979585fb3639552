#!/bin/bash
# round 2, call U: single-pick pairs drawn inside step1_cdf_kernel (GG_S1_DIRECT) A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
for v in main nodirect; do
  L=graphgan_b200/libgraphgan_b200.$v.so; [ $v = main ] && L=graphgan_b200/libgraphgan_b200.so
  GG_LIB=$L timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --g-steps 2 --verify 4 > $O/u_bench_$v.json 2> $O/u_bench_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/u_bench_$v.json").read().strip().splitlines()[-1])
    k=d["roofline"]["k1_stage"]
    print("$v", round(d["value"]/1e6,2), "M/s e2e", round(d["e2e"]["value"]/1e6,2), d["parity"]["mismatches"], d["ms_per_step"], k["hub_scores_root_cdf_ms"], k["root_step_step1_cdf_ms"], k["walk_kernel_ms"], k["finalize_emit_ms"], d["rates"]["g_mode"]["samples_per_s"])
except Exception as e:
    print("$v failed", e); print(open("$O/u_bench_$v.err").read()[-1500:])
PY
done
timeout 1200 python -m pytest tests/test_walk_gpu.py tests/test_config_parity_gpu.py tests/test_updates_gpu.py -q -m gpu -x > $O/u_pytest_main.log 2>&1
tail -n 2 $O/u_pytest_main.log
