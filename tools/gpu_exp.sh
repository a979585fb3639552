mkdir -p gpurun_out
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2> gpurun_out/bench_ref.time; grep real gpurun_out/bench_ref.time; cut -c1-160 gpurun_out/bench_ref.json; tail -c 700 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
( time timeout 1200 python bench.py --workload powerlaw_10m --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bench_10m.json 2> gpurun_out/bench_10m.err ) 2>&1 | grep real; cat gpurun_out/bench_10m.json | cut -c1-1500; tail -5 gpurun_out/bench_10m.err
