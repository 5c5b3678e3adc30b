mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/q_bench_n2.json 2> gpurun_out/q_bench_n2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/q_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/q_bench_n2.json').read().strip().splitlines()[-1])
for k in ('value','scaling','ms_per_step','e2e','p50_step_latency_ms','clocks','nccl_check','other_scaling','configs_multi_gpu','config'): print(k, d.get(k))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/q_ref_n2.json 2> gpurun_out/q_ref_n2.err; echo "ref n2 rc=$?"; cut -c1-400 gpurun_out/q_ref_n2.json
