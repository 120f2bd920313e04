for m in del keep; do
  timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tools/scratch/nccl_graph_teardown.py $m 2>&1 | grep -v "^W\|^\*\*\*" | tail -6
  echo "mode $m rc=$?"
done
start=$(date +%s)
timeout 200 python -m pytest tests/test_gpu_dp.py -m gpu -q -s -k product 2>&1 | grep -v "^E   " | tail -15 > gpurun_out/r02_t5_dp.txt; tail -6 gpurun_out/r02_t5_dp.txt
echo "dp test wall=$(( $(date +%s) - start )) s"
