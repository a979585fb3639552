mkdir -p gpurun_out
nproc; free -g | head -2; df -h /dev/shm | tail -1
( time timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2> gpurun_out/bench_ref.time; cat gpurun_out/bench_ref.time; cut -c1-200 gpurun_out/bench_ref.json; tail -c 500 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
timeout 900 python -m pytest tests/test_walk_gpu.py -q -m gpu -x -k "bfs" 2>&1 | tail -3
( time python bench.py --no-cpu-baseline > gpurun_out/b_cached.json 2> gpurun_out/b_cached.err ) 2>&1 | grep real; python -c "
import json; d=json.load(open('gpurun_out/b_cached.json')); print(d['value']/1e6, d['ms_per_step'], d['roofline']['traffic'], d['roofline']['frac'])"
