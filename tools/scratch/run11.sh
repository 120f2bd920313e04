timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02_t6_pytest.txt; tail -4 gpurun_out/r02_t6_pytest.txt
timeout 200 python tools/bench_samplers.py > gpurun_out/r02_samplers_hbm.txt 2>&1; cat gpurun_out/r02_samplers_hbm.txt | cut -c1-140
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_v3.json 2> gpurun_out/r02_bench_v3.err
python - <<EOF
import json
r=json.load(open("gpurun_out/r02_bench_v3.json")); t=r["train"]
print("render", round(r["value"]), "e2e", round(r["e2e"]["value"]), "share", r["roofline"]["kernel_share_of_step"], "mlp TF", r["roofline"]["achieved"], "train ms", t["ms_per_step"], "clk", r["clocks"], "launches", r["gpu_launches"], "traffic", r["roofline"]["traffic"])
EOF
