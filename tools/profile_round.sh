# Round profile: plain run first, then the ncu launch list of the same command, then one --set full capture per kernel.
# Summaries: tools/launch_shares.py, tools/summarize_ncu.py -> profiles/.
set -x
BC="python bench.py --steps 1 --warmup 3 --diffusion-steps 10 --no-cpu-baseline"
timeout 300 $BC > gpurun_out/r1f_bench_short.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r1f_launches.csv $BC > gpurun_out/r1f_ncu_list.log 2>&1
# -k matches the base name only (no template arguments): pick the launch by its index among the 30 conv_t3 launches
# of a step (tools/launch_shares.py prints the order): +12 = a C=512 k5 conv <64,1,2,2>, +8 = a C=256 k5 conv <32,1,2,2>
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_t3_kernel -s 162 -c 1 -f -o gpurun_out/r1f_t3_64 $BC > gpurun_out/r1f_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_t3_kernel -s 158 -c 1 -f -o gpurun_out/r1f_t3_32 $BC > gpurun_out/r1f_ncu2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"step_project_fused" -s 5 -c 1 -f -o gpurun_out/r1f_step $BC > gpurun_out/r1f_ncu3.log 2>&1
timeout 200 python tools/step_times.py 262144 > gpurun_out/r1f_stream.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"step_pointwise" -s 30 -c 1 -f -o gpurun_out/r1f_stream_inj python tools/step_times.py 262144 > gpurun_out/r1f_ncu4.log 2>&1
timeout 200 python tools/layer_times.py pointmaze 1 > gpurun_out/r1f_layers_b1.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_small" -s 60 -c 1 -f -o gpurun_out/r1f_small python tools/layer_times.py pointmaze 1 > gpurun_out/r1f_ncu5.log 2>&1
ls -la gpurun_out/r1f_*
