#!/bin/bash
# round 2, call R (2 GPUs): NCCL tests (root sharding, data-parallel updates, both transports), 2-GPU bench lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r_gpus.txt
timeout 900 python -m pytest tests/test_dist.py -q -m gpu > $O/r_pytest_dist_2gpu.log 2>&1
echo "dist rc=$?" >> $O/r_pytest_dist_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --verify 4 > $O/r_bench_2gpu.json 2> $O/r_bench_2gpu.err
for tp in nccl p2p; do
  timeout 600 $TR bench.py --gpus 2 --phase update --transport $tp --steps 50 --warmup 5 > $O/r_phase_update_2gpu_$tp.json 2> $O/r_phase_update_2gpu_$tp.err
done
tail -n 3 $O/r_pytest_dist_2gpu.log
head -c 400 $O/r_bench_2gpu.json; echo
cut -c 1-400 $O/r_phase_update_2gpu_*.json
tail -n 5 $O/r_bench_2gpu.err
