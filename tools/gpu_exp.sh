mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_updates_gpu.py -q -m gpu -x 2>&1 | tail -2
timeout 200 python tools/bench_updates.py 1000000 128 64 300 2>&1 | grep -v "CTA 0" | tail -6
timeout 200 python tools/bench_updates.py 2>&1 | grep -v "CTA 0" | tail -6
