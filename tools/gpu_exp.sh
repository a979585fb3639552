mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_updates_gpu.py -q -m gpu -x 2>&1 | tail -3
timeout 300 python tools/bench_updates.py 2>&1 | tail -8
timeout 600 python -m pytest tests/test_walk_gpu.py -q -m gpu -x 2>&1 | tail -3
for v in "new:" "t256:--hub-threshold 256" "r4k:--roots 4096"; do
  name=${v%%:*}; flags=${v#*:}
  timeout 600 python bench.py --no-cpu-baseline $flags > gpurun_out/b_$name.json 2> gpurun_out/b_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/b_$name.json"))
    r=d["roofline"]
    print("$name", "value %.2fM e2e %.2fM ms %.3f walk %.3f pre %.3f"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"],r["walk_kernel_ms"],r["precompute_ms"]), d["walk"]["warp_cycle_share"], d["config"]["workload"][-90:-50])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/b_$name.err").read()[-1500:])
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:_kernel -c 400 --csv --log-file gpurun_out/launches_d1.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_d1.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_d1.csv")) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if "gg::" not in r[ki]: continue
    v = float(r[vi].replace(",", "")); v = v / 1e6 if r[ui] in ("ns", "nsecond") else (v / 1e3 if r[ui].startswith("u") else v)
    a = agg.setdefault(r[ki][:60], [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, t) in agg.items(): print("%-62s %4d launches %10.3f ms total %8.3f ms each" % (k, n, t, t / n))
PY
