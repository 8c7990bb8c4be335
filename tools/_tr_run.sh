python -m pytest tests/test_gpu_treasure.py -q -x 2>&1 | grep -v "^  \|Warning" | tail -40 > gpurun_out/s2_treasure.log; grep -n "^E\|passed\|failed" gpurun_out/s2_treasure.log | head -8
python tools/bench_env_configs.py 2>&1 | tail -1
ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_treasure -s 3 -c 1 python tools/profile_target.py treasure 2>&1 | grep -E "inst_executed|issue_active|time_duration|warps_active"
