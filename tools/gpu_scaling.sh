# N-GPU bench line (strong scaling by default, weak beside it, NCCL bit-exactness check, configs 3/4 as sharded legs).
# usage: gpurun --gpus N -- 'bash tools/gpu_scaling.sh N'
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench n$N rc=$?"; tail -3 gpurun_out/bench_n$N.err
