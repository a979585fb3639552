mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_walk_gpu.py -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --steps 30 > gpurun_out/bench_walk.json 2> gpurun_out/bench_walk.err
python -c "
import json,sys; d=json.load(open('gpurun_out/bench_walk.json')); print({k:round(d[k],3) for k in ['value','ms_per_step']}, round(d['e2e']['value']), {k:round(d['roofline'][k],3) for k in ['precompute_ms','walk_kernel_ms']}, 'bfs', d['walk']['bfs_build_s'])"
tail -2 gpurun_out/bench_walk.err
