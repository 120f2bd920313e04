nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q -s 2>&1 | tail -40 > gpurun_out/r02_t4_dp.txt; tail -15 gpurun_out/r02_t4_dp.txt
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_2gpu_v1.json 2> gpurun_out/r02_bench_2gpu_v1.err
tail -5 gpurun_out/r02_bench_2gpu_v1.err | cut -c1-300
grep -c "NCCL INFO" gpurun_out/r02_bench_2gpu_v1.err
python - <<EOF
import json
try:
    r=json.load(open("gpurun_out/r02_bench_2gpu_v1.json")); t=r["train"]
    print("render", round(r["value"]), "e2e", round(r["e2e"]["value"]), "strong", r.get("strong"), "train ms", t["ms_per_step"], "rays/s", round(t["value"]), "loss", t["loss_first"], t["loss_last"])
except Exception as e: print("ERR", e)
EOF
