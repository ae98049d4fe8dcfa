mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_gpu_big_configs.py -q -k "multi or sharded or two" 2>&1 | tail -4 > gpurun_out/r2q_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2q_bench_n2.json 2> gpurun_out/r2q_bench_n2.err
W=$(mktemp -d); cd $W
python $GRAFT_REPO_ROOT/tools/gen_config4_phylip.py 4 > /dev/null 2>&1
( time PLK_GPUS=2 PLK_SHIM_VERBOSE=1 timeout 400 $GRAFT_REPO_ROOT/integration/_build/phyml_b200 -i c4.phy -u c4.nwk -d nt -m GTR -c 4 -a 0.5 -f 0.30,0.20,0.25,0.25 -o n -b 0 --r_seed 1 --no_memory_check ) > $GRAFT_REPO_ROOT/gpurun_out/r2q_c4_dropin_2gpu.log 2>&1
