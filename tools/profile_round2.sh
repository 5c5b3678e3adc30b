# Round-2 profile: plain run first, then the ncu launch list of the same command, then --set full captures.
# The reports are summarised ON the GPU box (tools/summarize_ncu.py --out gpurun_out) and deleted: only gpurun_out/
# travels back and it is capped at 64 MiB.  Copy gpurun_out/r02_* and ncu_traffic.json to profiles/ afterwards.
set -x
mkdir -p gpurun_out
O=gpurun_out
BC="python bench.py --steps 1 --warmup 3 --diffusion-steps 10 --no-cpu-baseline --no-extra-legs"
timeout 300 $BC > $O/r02_bench_short.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --print-kernel-base demangled -c 1500 --csv --log-file $O/r02_launches_bench_pointmaze.csv $BC > $O/r02_ncu_list.log 2>&1
python tools/launch_shares.py $O/r02_launches_bench_pointmaze.csv $O/r02_launch_shares.md "$BC" > /dev/null
# one U-Net pass = 6 conv_chain launches (5 / 5 / 9 / 5 / 5 / 1 convs): skip the warm-up plans, capture one pass
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_chain_kernel -s 120 -c 6 -f -o /tmp/r02_chain $BC > $O/r02_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"step_project_fused" -s 25 -c 1 -f -o /tmp/r02_step $BC > $O/r02_ncu2.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"conv_tc_kernel" -s 140 -c 7 -f -o /tmp/r02_tc $BC > $O/r02_ncu3.log 2>&1
python tools/summarize_ncu.py --out $O r02 pointmaze conv_chain_kernel /tmp/r02_chain.ncu-rep step_project_fused_kernel /tmp/r02_step.ncu-rep conv_tc_kernel /tmp/r02_tc.ncu-rep > /dev/null
# stall samples of the narrow (L=32) chain and of the dominant 9-conv chain
ncu -i /tmp/r02_chain.ncu-rep --page source --csv --kernel-id :::1 > /tmp/src1.csv 2>/dev/null; python tools/ncu_stalls.py /tmp/src1.csv 25 > $O/r02_stalls_chain_L32.txt 2>&1
ncu -i /tmp/r02_chain.ncu-rep --page source --csv --kernel-id :::3 > /tmp/src3.csv 2>/dev/null; python tools/ncu_stalls.py /tmp/src3.csv 25 > $O/r02_stalls_chain_C512.txt 2>&1
timeout 200 python tools/step_times.py 262144 > $O/r02_stream.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"step_pointwise" -s 30 -c 1 -f -o /tmp/r02_stream_inj python tools/step_times.py 262144 > $O/r02_ncu4.log 2>&1
python tools/summarize_ncu.py --out $O r02 stream_b262144 step_pointwise_kernel /tmp/r02_stream_inj.ncu-rep > /dev/null
# the same kernel with in-kernel Philox noise (the LEAN, software-pipelined instantiation): first launches of step_times.py
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"step_pointwise" -s 5 -c 1 -f -o /tmp/r02_stream_philox python tools/step_times.py 262144 > $O/r02_ncu4b.log 2>&1
python tools/summarize_ncu.py --out $O r02 stream_philox_b262144 step_pointwise_kernel /tmp/r02_stream_philox.ncu-rep > /dev/null
# HalfCheetah: the GroupNorm-width-256 chain (9 convs, C_out 2048)
cat > /tmp/prof_hc.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import bench
dev = torch.device('cuda', 0)
w = dict(bench.WORKLOADS['halfcheetah']); w['S'] = 20
net, dif = bench.build_policy(w, 1024, 'bf16', dev, latency_max_batch=0)
eng = dif.engine(32, dev)
x = torch.randn(1024, 32, 23, device=dev)
for _ in range(3):
    eng.unet_forward(x, step=3)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none -k regex:conv_chain_kernel -s 12 -c 6 -f -o /tmp/r02_chain_hc python /tmp/prof_hc.py > $O/r02_ncu5.log 2>&1
python tools/summarize_ncu.py --out $O r02 halfcheetah conv_chain_kernel /tmp/r02_chain_hc.ncu-rep > /dev/null
# B = 1 (get_action): the latency kernels
timeout 400 ncu --set full --clock-control none -k regex:"conv_small" -s 60 -c 1 -f -o /tmp/r02_small python tools/layer_times.py pointmaze 1 > $O/r02_ncu6.log 2>&1
python tools/summarize_ncu.py --out $O r02 pointmaze_b1 conv_small_kernel /tmp/r02_small.ncu-rep > /dev/null
# the fp32 sibling's kernels (ill-conditioned leading step): TF32 mma.sync conv and the warp GroupNorm+Mish
# (the sibling's convs are conv_tc_kernel<BN, GW, true>: the first conv_tc launches of a plan; bf16 conv_tc launches follow)
timeout 400 ncu --set full --clock-control none -k regex:"conv_tc_kernel" -s 111 -c 12 -f -o /tmp/r02_tf32 $BC > $O/r02_ncu7.log 2>&1
python tools/summarize_ncu.py --out $O r02 pointmaze_sibling conv_tc_kernel_tf32 /tmp/r02_tf32.ncu-rep > /dev/null
du -sh $O; ls -la $O | head -40
