start=$(date +%s)
timeout 240 python -m pytest tests/test_gpu_dp.py -m gpu -q -s 2>&1 | grep -v "^E   " | tail -15 > gpurun_out/r02_t5_dp.txt; tail -6 gpurun_out/r02_t5_dp.txt
echo "dp tests wall=$(( $(date +%s) - start )) s"
start=$(date +%s)
NCCL_DEBUG=INFO timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02_bench_2gpu_v3.json 2> gpurun_out/r02_bench_2gpu_v3.err
echo "bench rc=$? wall=$(( $(date +%s) - start )) s"
