#!/bin/bash
# round 2, call Z2: TMA-staged hub enumeration A/B on the final kernels (bench --no-tma), same box back to back
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --g-steps 2 --verify 4 > $O/z2_bench_tma.json 2> $O/z2_bench_tma.err
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --g-steps 2 --verify 4 --no-tma > $O/z2_bench_notma.json 2> $O/z2_bench_notma.err
for v in tma notma; do python - <<PY
import json
d=json.loads(open("$O/z2_bench_$v.json").read().strip().splitlines()[-1]); k=d["roofline"]["k1_stage"]
print("$v", round(d["value"]/1e6,2), "M/s e2e", round(d["e2e"]["value"]/1e6,2), d["parity"]["mismatches"], d["ms_per_step"], k["hub_scores_root_cdf_ms"], k["root_step_step1_cdf_ms"], k["walk_kernel_ms"])
PY
done
