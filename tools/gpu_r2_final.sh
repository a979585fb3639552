#!/bin/bash
# round 2, final 1-GPU call: whole GPU test suite, smoke, every bench line quoted in DESIGN.md / profiles/README.md, ncu captures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
python -c "from graphgan_b200 import _build; print(_build.source_hash())" > $O/z_source_hash.txt
timeout 1500 python -m pytest tests -q -m gpu > $O/z_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/z_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/z_smoke.log 2>&1
echo "smoke rc=$?" >> $O/z_smoke.log
timeout 900 python bench.py > $O/z_bench.json 2> $O/z_bench.err
echo "bench rc=$?" >> $O/z_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/z_bench_reference.json 2> $O/z_bench_reference.err
timeout 600 python bench.py --impl reference --workload c1_cagrqc --score-mode literal --steps 1 --warmup 0 --cpu-seconds 20 > $O/z_ref_c1_literal.json 2> $O/z_ref_c1_literal.err
for ph in bfs reward adam update; do
  timeout 600 python bench.py --phase $ph --steps 20 --warmup 3 > $O/z_phase_$ph.json 2> $O/z_phase_$ph.err
done
timeout 900 python bench.py --workload er_100k --no-cpu-baseline > $O/z_bench_er_100k.json 2> $O/z_bench_er_100k.err
timeout 900 python bench.py --workload c1_cagrqc --no-cpu-baseline --verify 0 > $O/z_bench_c1.json 2> $O/z_bench_c1.err
BENCH1="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --verify 0 --g-steps 0"
FAM="hub_score_kernel|root_cdf_kernel|root_step_kernel|step1_cdf_kernel|walk_kernel|flat_start_kernel|flat_enum_kernel|flat_choose_kernel"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/z_launches.csv $BENCH1 > $O/z_ncu1.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:"$FAM" -s 14 -c 28 -o $O/z_k1_metrics -f $BENCH1 > $O/z_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"flat_choose_kernel" -s 5 -c 1 -o $O/z_prof_choose -f $BENCH1 > $O/z_ncu3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"step1_cdf_kernel" -s 1 -c 1 -o $O/z_prof_step1 -f $BENCH1 > $O/z_ncu4.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"bfs_kernel" -s 1 -c 1 -o $O/z_prof_bfs -f \
    python bench.py --phase bfs --bfs-roots 296 --steps 1 --warmup 1 > $O/z_ncu5.log 2>&1
tail -n 3 $O/z_pytest_gpu.log $O/z_smoke.log $O/z_bench.err
head -c 600 $O/z_bench.json; echo
cut -c 1-200 $O/z_phase_*.json
