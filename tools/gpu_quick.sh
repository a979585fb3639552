mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -8
for F in "" "--depth1"; do
timeout 300 python bench.py --no-cpu-baseline --steps 30 $F > gpurun_out/bench_walk.json 2> gpurun_out/bench_walk.err
python -c "
import json,sys; d=json.load(open('gpurun_out/bench_walk.json')); print('$F', {k:round(d[k],3) for k in ['value','ms_per_step']}, round(d['e2e']['value']), {k:round(d['roofline'][k],3) for k in ['precompute_ms','walk_kernel_ms']}, d['walk']['warp_cycle_share'])"
tail -2 gpurun_out/bench_walk.err
done
