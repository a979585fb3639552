mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; wc -l gpurun_out/bench_final.json; cut -c1-200 gpurun_out/bench_final.json
