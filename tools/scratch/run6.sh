timeout 300 python -m pytest tests/test_gpu_dp.py -m gpu -q -s -k product 2>&1 | grep -v "^E   " | tail -15 > gpurun_out/r02_t5_dp.txt; tail -6 gpurun_out/r02_t5_dp.txt
start=$(date +%s)
NCCL_DEBUG=INFO timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02_bench_2gpu_v2.json 2> gpurun_out/r02_bench_2gpu_v2.err
echo "bench rc=$? wall=$(( $(date +%s) - start )) s"
python - <<EOF
import json
try:
    r=json.load(open("gpurun_out/r02_bench_2gpu_v2.json")); t=r["train"]
    print("render", round(r["value"]), "e2e", round(r["e2e"]["value"]), "strong ms", r["strong"]["ms_per_frame"], "train ms", t["ms_per_step"], "loss", t["loss_first"], t["loss_last"])
except Exception as e: print("ERR", e)
EOF
