mkdir -p gpurun_out
python tools/dbg_updates.py 2>&1 | tail -30
python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
