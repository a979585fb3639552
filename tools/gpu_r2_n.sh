#!/bin/bash
# round 2, call N: thread-per-walk finalize / emit kernels, S1_MIN_WALKS = 1 variant
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_walk_gpu.py tests/test_config_parity_gpu.py tests/test_updates_gpu.py -q -m gpu -x > $O/n_pytest.log 2>&1
echo "pytest rc=$?" >> $O/n_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --g-steps 2 --verify 4 > $O/n_bench.json 2> $O/n_bench.err
GG_LIB=graphgan_b200/libgraphgan_b200.s1min1.so timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --g-steps 2 --verify 4 > $O/n_bench_s1min1.json 2> $O/n_bench_s1min1.err
GG_LIB=graphgan_b200/libgraphgan_b200.s1min1.so timeout 900 python -m pytest tests/test_walk_gpu.py -q -m gpu -x > $O/n_pytest_s1min1.log 2>&1
tail -n 3 $O/n_pytest.log $O/n_pytest_s1min1.log
for f in n_bench n_bench_s1min1; do python - <<PY
import json
try:
    d=json.loads(open("$O/$f.json").read().strip().splitlines()[-1])
    k=d["roofline"]["k1_stage"]
    print("$f", round(d["value"]/1e6,2), "M/s e2e", round(d["e2e"]["value"]/1e6,2), d["parity"]["mismatches"], d["ms_per_step"], k["hub_scores_root_cdf_ms"], k["root_step_step1_cdf_ms"], k["walk_kernel_ms"], k["finalize_emit_ms"], d["rates"]["g_mode"]["samples_per_s"])
except Exception as e:
    print("$f failed", e); print(open("$O/$f.err").read()[-1500:])
PY
done
