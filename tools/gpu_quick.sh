mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -5
for C in 0 0; do
python bench.py --no-cpu-baseline --chunk-walks $C --steps 30 > gpurun_out/bench_C$C.json 2> gpurun_out/bench_C$C.err
python -c "
import json,sys; d=json.load(open('gpurun_out/bench_C$C.json')); print('C=$C', {k:round(d[k],3) for k in ['value','ms_per_step']}, round(d['e2e']['value']), {k:round(d['roofline'][k],3) for k in ['precompute_ms','walk_kernel_ms']}, d['walk']['warp_cycle_share'])"
tail -2 gpurun_out/bench_C$C.err
done
