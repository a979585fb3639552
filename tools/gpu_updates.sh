mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_updates_gpu.py tests/test_dist.py -q -m gpu -x 2>&1 | tail -5
timeout 300 python tools/bench_updates.py 2>&1 | tail -12
timeout 300 python tools/bench_updates.py 1000000 128 64 300 2>&1 | tail -8
timeout 300 python tools/bench_updates.py 5242 50 1024 5000 2>&1 | tail -12
timeout 600 python -m pytest tests/test_walk_gpu.py -q -m gpu -x 2>&1 | tail -3
for v in "d1q:--depth1" "d1static:--depth1 --file-order" "base:"; do
  name=${v%%:*}; flags=${v#*:}
  timeout 600 python bench.py --no-cpu-baseline $flags > gpurun_out/b_$name.json 2> gpurun_out/b_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/b_$name.json"))
    r=d["roofline"]
    print("$name", "value %.2fM e2e %.2fM ms %.3f walk %.3f pre %.3f"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"],r["walk_kernel_ms"],r["precompute_ms"]), d["walk"]["warp_cycle_share"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/b_$name.err").read()[-1500:])
PY
done
