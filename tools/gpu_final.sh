mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err; wc -l gpurun_out/bench_1m.json; cut -c1-330 gpurun_out/bench_1m.json; tail -2 gpurun_out/bench_1m.err
timeout 300 ncu --set full --clock-control none -k regex:adam_kernel -s 5 -c 1 -o gpurun_out/prof_adam_1m python tools/prof_updates.py steps 12 1000000 128 > gpurun_out/ncu_adam.log 2>&1; tail -1 gpurun_out/ncu_adam.log
