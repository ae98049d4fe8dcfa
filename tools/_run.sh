mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2c_gpus.txt
python -m pytest tests/test_gpu_multi.py tests/test_gpu_big_configs.py -m gpu -q -k "multi or sharded or two" 2>&1 | tail -40 > gpurun_out/r2c_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err
cd tests/golden
ARGS="-i synth_dna_deep.phy -u synth_dna_deep.nwk -d nt -m GTR -c 4 -a 0.5 -f e -o n -b 0 --r_seed 1 --no_memory_check"
PLK_SHIM_VERBOSE=1 ../../integration/_build/phyml_b200 $ARGS > ../../gpurun_out/r2c_dropin_1gpu.log 2>&1
PLK_GPUS=2 PLK_SHIM_VERBOSE=1 ../../integration/_build/phyml_b200 $ARGS > ../../gpurun_out/r2c_dropin_2gpu.log 2>&1
rm -f synth_dna_deep.phy_phyml_*
