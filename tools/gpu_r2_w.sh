#!/bin/bash
# round 2, call W: BASELINE configs[4] shape on one GPU (N = 10M, avg-deg 8, n_emb 256), final kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python bench.py --workload powerlaw_10m --steps 10 --warmup 3 --no-cpu-baseline --verify 3 --g-steps 1 > gpurun_out/w_bench_powerlaw_10m.json 2> gpurun_out/w_bench_powerlaw_10m.err
echo "rc=$?"; tail -n 3 gpurun_out/w_bench_powerlaw_10m.err; head -c 700 gpurun_out/w_bench_powerlaw_10m.json
