#!/bin/bash
# round 2, call M: flat steps with walk records (default flat_steps = 4): full GPU suite, bench A/B, launch list, ncu traffic + ncu of flat_choose
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
python -c "from graphgan_b200 import _build; print(_build.source_hash())" > $O/m_source_hash.txt
timeout 1500 python -m pytest tests -q -m gpu > $O/m_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/m_pytest_gpu.log
for f in 0 3 4 5; do
  timeout 600 python bench.py --flat-steps $f --steps 20 --warmup 3 --no-cpu-baseline --g-steps 2 --verify 4 > $O/m_bench_flat$f.json 2> $O/m_bench_flat$f.err
done
timeout 600 python bench.py --phase bfs --steps 5 --warmup 2 > $O/m_phase_bfs.json 2> $O/m_phase_bfs.err
BENCH1="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --verify 0 --g-steps 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/m_launches.csv $BENCH1 > $O/m_ncu1.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:"hub_score_kernel|root_cdf_kernel|root_step_kernel|step1_cdf_kernel|walk_kernel|flat_start_kernel|flat_enum_kernel|flat_choose_kernel" -s 14 -c 28 -o $O/m_k1_metrics -f $BENCH1 > $O/m_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"flat_choose_kernel" -s 5 -c 1 -o $O/m_prof_choose -f $BENCH1 > $O/m_ncu3.log 2>&1
tail -n 3 $O/m_pytest_gpu.log
for f in 0 3 4 5; do python - <<PY
import json
try:
    d=json.loads(open("$O/m_bench_flat$f.json").read().strip().splitlines()[-1])
    print("flat $f", round(d["value"]/1e6,2), "M/s e2e", round(d["e2e"]["value"]/1e6,2), d["parity"]["mismatches"], d["roofline"]["k1_stage"]["walk_kernel_ms"], d["rates"]["g_mode"]["samples_per_s"])
except Exception as e:
    print("flat $f failed", e); print(open("$O/m_bench_flat$f.err").read()[-1500:])
PY
done
grep -o '"ms_per_root": [0-9.]*' $O/m_phase_bfs.json
