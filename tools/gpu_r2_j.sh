#!/bin/bash
# round 2, call J (re-entry): GPU test suite + smoke + default bench + phase lines + launch list + walk-kernel ncu, v7 sources
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
python -c "from graphgan_b200 import _build; print(_build.source_hash())" > $O/j_source_hash.txt
timeout 1500 python -m pytest tests -q -m gpu > $O/j_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/j_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/j_smoke.log 2>&1
echo "smoke rc=$?" >> $O/j_smoke.log
timeout 900 python bench.py > $O/j_bench.json 2> $O/j_bench.err
echo "bench rc=$?" >> $O/j_bench.err
for ph in bfs reward adam update; do
  timeout 600 python bench.py --phase $ph --steps 20 --warmup 3 > $O/j_phase_$ph.json 2> $O/j_phase_$ph.err
done
BENCH1="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --verify 0 --g-steps 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/j_launches.csv $BENCH1 > $O/j_ncu1.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:"hub_score_kernel|root_cdf_kernel|root_step_kernel|step1_cdf_kernel|walk_kernel" -s 5 -c 10 -o $O/j_k1_metrics -f $BENCH1 > $O/j_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"walk_kernel" -s 1 -c 1 -o $O/j_prof_walk -f $BENCH1 > $O/j_ncu3.log 2>&1
tail -n 3 $O/j_pytest_gpu.log $O/j_smoke.log $O/j_bench.err
head -c 600 $O/j_bench.json; echo
cut -c 1-200 $O/j_phase_*.json
