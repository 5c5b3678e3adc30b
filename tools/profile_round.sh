set -x
BC="python bench.py --steps 1 --warmup 3 --diffusion-steps 10 --no-cpu-baseline"
timeout 300 $BC > gpurun_out/r1f_bench_short.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r1f_launches.csv $BC > gpurun_out/r1f_ncu_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_t3_kernel<64" -s 40 -c 1 -f -o gpurun_out/r1f_t3_64 $BC > gpurun_out/r1f_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_t3_kernel<32" -s 40 -c 1 -f -o gpurun_out/r1f_t3_32 $BC > gpurun_out/r1f_ncu2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"step_project_fused" -s 5 -c 1 -f -o gpurun_out/r1f_step $BC > gpurun_out/r1f_ncu3.log 2>&1
timeout 200 python tools/step_times.py 262144 > gpurun_out/r1f_stream.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"step_pointwise" -s 30 -c 1 -f -o gpurun_out/r1f_stream_inj python tools/step_times.py 262144 > gpurun_out/r1f_ncu4.log 2>&1
timeout 200 python tools/layer_times.py pointmaze 1 > gpurun_out/r1f_layers_b1.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_small" -s 60 -c 1 -f -o gpurun_out/r1f_small python tools/layer_times.py pointmaze 1 > gpurun_out/r1f_ncu5.log 2>&1
ls -la gpurun_out/r1f_*
