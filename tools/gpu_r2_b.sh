#!/bin/bash
# round 2, call B: BFS v2 + f4 + checkpoint tests, full bench line, phase lines, ncu captures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_walk_gpu.py tests/test_config_parity_gpu.py -x -q -m gpu > $O/b_walk.log 2>&1
echo "walk rc=$?" >> $O/b_walk.log
timeout 900 python -m pytest tests/test_updates_gpu.py -x -q -m gpu -s > $O/b_updates.log 2>&1
echo "updates rc=$?" >> $O/b_updates.log
timeout 900 python bench.py > $O/b_bench.json 2> $O/b_bench.err
echo "bench rc=$?" >> $O/b_bench.err
for ph in bfs reward adam; do
  timeout 600 python bench.py --phase $ph --steps 10 --warmup 3 > $O/b_phase_$ph.json 2> $O/b_phase_$ph.err
  echo "phase $ph rc=$?" >> $O/b_phase_$ph.err
done
# ncu: (1) launch list of the default bench command, (2) DRAM bytes of the K1 kernels, (3) full sets of the top kernels
BENCH1="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --verify 0 --g-steps 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/b_launches.csv $BENCH1 > $O/b_ncu1.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:"hub_score_kernel|root_cdf_kernel|root_step_kernel|step1_cdf_kernel|walk_kernel" -s 5 -c 10 -o $O/b_k1_metrics -f $BENCH1 > $O/b_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"walk_kernel" -s 1 -c 1 -o $O/b_prof_walk -f $BENCH1 > $O/b_ncu3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"step1_cdf_kernel" -s 1 -c 1 -o $O/b_prof_step1 -f $BENCH1 > $O/b_ncu4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bfs_kernel" -s 1 -c 1 -o $O/b_prof_bfs -f \
    python bench.py --phase bfs --bfs-roots 296 --steps 1 --warmup 1 > $O/b_ncu5.log 2>&1
tail -n 3 $O/b_walk.log $O/b_updates.log $O/b_bench.err
head -c 1200 $O/b_bench.json; echo
cat $O/b_phase_*.json | cut -c 1-400
