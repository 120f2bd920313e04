B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train"
$B > gpurun_out/r02_plain_render.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_render.csv $B > gpurun_out/r02_ncu_render_list.log 2>&1
echo "list rc=$?"
$B > gpurun_out/r02_plain_render.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mlp_tc3 -s 1 -c 1 -o gpurun_out/r02_mlp_tc3_frame -f $B > gpurun_out/r02_ncu_render_full.log 2>&1
echo "full rc=$?"
T="python tools/ncu_train_step.py"
$T > gpurun_out/r02_plain_train.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'mlp_tc3|bwd3|wgrad' -s 12 -c 6 -o gpurun_out/r02_train_kernels -f $T > gpurun_out/r02_ncu_train_full.log 2>&1
echo "train rc=$?"
ls -la gpurun_out/*.ncu-rep
