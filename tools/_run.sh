mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_big_configs.py -x -q -k "aa or proteic or lg or synthetic or AA" 2>&1 | tail -3 > gpurun_out/r2l_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --workload aa_200x50k > gpurun_out/r2l_aa3.json 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_traverse_aa' -s 2 -c 1 -f -o gpurun_out/prof_aa3c python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --workload aa_200x50k > gpurun_out/ncu_aa3c.log 2>&1
