mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/y_bench_n$N.json 2> gpurun_out/y_bench_n$N.err; echo "bench n$N rc=$?"; tail -3 gpurun_out/y_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/y_bench_n$N.json').read().strip().splitlines()[-1])
for k in ('value','scaling','ms_per_step','e2e','p50_step_latency_ms','clocks','nccl_check','other_scaling'): print(k, d.get(k))
for c in d.get('configs_multi_gpu', []): print({k: c.get(k) for k in ('workload','B_total','B_per_gpu','p50_step_latency_ms_max_over_ranks','plans_per_s_extrapolated','error')})
PY
