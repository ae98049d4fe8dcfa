mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --workload aa_200x50k > gpurun_out/r2g_aa2.json 2>&1
PLK_AA_V1=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --workload aa_200x50k > gpurun_out/r2g_aa1.json 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_traverse_aa' -s 2 -c 1 -f -o gpurun_out/prof_aa2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --workload aa_200x50k > gpurun_out/ncu_aa2.log 2>&1
