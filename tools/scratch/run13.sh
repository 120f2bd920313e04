timeout 500 python -m pytest tests/test_gpu_fused_composite.py tests/test_gpu_training.py tests/test_gpu_trajectory.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02_t8_pytest.txt; tail -4 gpurun_out/r02_t8_pytest.txt
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_v5.json 2> gpurun_out/r02_bench_v5.err
python - <<EOF
import json
r=json.load(open("gpurun_out/r02_bench_v5.json")); t=r["train"]
print("render", round(r["value"]), "e2e", round(r["e2e"]["value"]), "train ms", t["ms_per_step"], "loss", t["loss_last"], "clk", r["clocks"]["sm_mhz"])
for k in t["roofline_train"]["kernels"]: print(k["kernel"], round(k["ms_per_step"],3), round(k["hbm_gbs"]), round(k["tflops"]))
EOF
