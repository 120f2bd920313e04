start=$(date +%s)
NCCL_DEBUG=INFO timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02_bench_8gpu_v1.json 2> gpurun_out/r02_bench_8gpu_v1.err
echo "bench rc=$? wall=$(( $(date +%s) - start )) s"
python - <<EOF
import json
try:
    r=json.load(open("gpurun_out/r02_bench_8gpu_v1.json")); t=r["train"]
    print("render", round(r["value"]), "e2e", round(r["e2e"]["value"]), "strong ms", r["strong"]["ms_per_frame"], r["strong"]["speedup_vs_one_rank_e2e"], "train ms", t["ms_per_step"], round(t["value"]), "loss", t["loss_first"], t["loss_last"], r["clocks"])
except Exception as e: print("ERR", e)
EOF
grep -c "NCCL INFO" gpurun_out/r02_bench_8gpu_v1.err
