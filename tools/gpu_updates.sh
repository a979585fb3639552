mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_updates_gpu.py -q -m gpu 2>&1 | tail -5
timeout 300 python tools/bench_updates.py 2>&1 | tail -8
timeout 300 python tools/bench_updates.py 1000000 128 64 300 2>&1 | tail -8
timeout 900 python tools/train_c1.py 3 2>&1 | grep -v "^reading\|^construct\|^building\|^epoch\|^start\|^training" | tail -12
