mkdir -p gpurun_out
nvidia-smi -L | head -4
python -m pytest tests/test_dist.py -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_dist.log
tail -8 gpurun_out/pytest_dist.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_1m_n2.json 2> gpurun_out/bench_1m_n2.err
cat gpurun_out/bench_1m_n2.json | cut -c1-700; tail -5 gpurun_out/bench_1m_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_n2.json; tail -2 gpurun_out/bench_ref_n2.err
