mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
python bench.py > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_1m.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 1 -c 1 -o gpurun_out/prof_walk python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hub_score -s 1 -c 1 -o gpurun_out/prof_hub python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_hub.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench_1m.json; tail -3 gpurun_out/bench_1m.err
