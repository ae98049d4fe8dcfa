mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2h_gpus.txt
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2h_bench_n$N.json 2> gpurun_out/r2h_bench_n$N.err
done
PLK_COMM=nccl python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2h_bench_n8_nccl.json 2> gpurun_out/r2h_bench_n8_nccl.err
python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -5 > gpurun_out/r2h_pytest_multi.log
