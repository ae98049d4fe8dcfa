mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_big_configs.py -x -q -k "aa or proteic or lg or synthetic or AA" 2>&1 | tail -3 > gpurun_out/r2p_pytest.log
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --workload aa_200x50k > gpurun_out/r2p_aa.json 2>&1
