#!/bin/bash
# round 2, call C (2 GPUs): NCCL-inside-the-C-ABI data-parallel step: parity test, update-phase bench at 1 and 2 GPUs,
# sampling bench at 2 GPUs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > $O/c_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_dist.py -x -q -m gpu > $O/c_dist.log 2>&1
echo "dist rc=$?" >> $O/c_dist.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python bench.py --phase update --steps 30 --warmup 5 > $O/c_update_1gpu.json 2> $O/c_update_1gpu.err
timeout 600 $TR --nproc-per-node 2 --master-port 29711 bench.py --gpus 2 --phase update --steps 30 --warmup 5 > $O/c_update_2gpu.json 2> $O/c_update_2gpu.err
echo "update2 rc=$?" >> $O/c_update_2gpu.err
timeout 600 $TR --nproc-per-node 2 --master-port 29713 bench.py --gpus 2 --phase update --transport p2p --steps 30 --warmup 5 > $O/c_update_2gpu_p2p.json 2> $O/c_update_2gpu_p2p.err
echo "update2 p2p rc=$?" >> $O/c_update_2gpu_p2p.err
timeout 600 python bench.py --phase adam --adam-path ldg --steps 20 --warmup 3 > $O/c_adam_ldg.json 2> $O/c_adam_ldg.err
timeout 600 python bench.py --phase adam --adam-path tma --steps 20 --warmup 3 > $O/c_adam_tma.json 2> $O/c_adam_tma.err
timeout 900 $TR --nproc-per-node 2 --master-port 29712 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --verify 4 > $O/c_bench_2gpu.json 2> $O/c_bench_2gpu.err
echo "bench2 rc=$?" >> $O/c_bench_2gpu.err
tail -n 4 $O/c_dist.log
cut -c 1-600 $O/c_update_1gpu.json $O/c_update_2gpu.json $O/c_update_2gpu_p2p.json $O/c_adam_ldg.json $O/c_adam_tma.json
cut -c 1-300 $O/c_bench_2gpu.json
tail -n 3 $O/c_update_2gpu.err $O/c_update_2gpu_p2p.err $O/c_bench_2gpu.err; tail -n 30 $O/c_dist.log
