mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_updates_gpu.py -q -m gpu -x 2>&1 | tail -3
timeout 300 python tools/bench_updates.py 2>&1 | tail -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:train_fused -c 1 -o gpurun_out/prof_fused python tools/prof_updates.py fused 1000 > gpurun_out/ncu_fused.log 2>&1; tail -2 gpurun_out/ncu_fused.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"adam_kernel|pair_grad_kernel" -s 20 -c 2 -o gpurun_out/prof_steps python tools/prof_updates.py steps 50 > gpurun_out/ncu_steps.log 2>&1; tail -2 gpurun_out/ncu_steps.log
