timeout 500 python -m pytest tests -m gpu -q -s 2>&1 | tail -60 > gpurun_out/r02_t3_pytest.txt; tail -6 gpurun_out/r02_t3_pytest.txt
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_v2.json 2> gpurun_out/r02_bench_v2.err; tail -3 gpurun_out/r02_bench_v2.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --eager-train > gpurun_out/r02_bench_v2_eager.json 2>/dev/null
python - <<EOF
import json
for f in ("gpurun_out/r02_bench_v2.json","gpurun_out/r02_bench_v2_eager.json"):
    try:
        r=json.load(open(f)); t=r["train"]; print(f, "render", round(r["value"]), "e2e", round(r["e2e"]["value"]), "train ms", t["ms_per_step"], "rays/s", round(t["value"]), "loss", t["loss_first"], t["loss_last"], "clk", r["clocks"]["sm_mhz"])
    except Exception as e: print(f, "ERR", e)
EOF
