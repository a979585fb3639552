mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
python bench.py > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:_kernel -c 400 --csv --log-file gpurun_out/launches_1m.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
for k in walk_kernel step1_cdf hub_score root_cdf; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/prof_$k python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$k.log 2>&1
done
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench_1m.json; tail -3 gpurun_out/bench_1m.err; cat gpurun_out/bench_ref.json; tail -2 gpurun_out/bench_ref.err
