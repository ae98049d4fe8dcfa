mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2f_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
PLK_NO_FUSED_EDGE=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/r2f_nofuse.json 2>&1
