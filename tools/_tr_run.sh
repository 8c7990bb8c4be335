python tools/bench_env_configs.py 2>&1 | tail -1
