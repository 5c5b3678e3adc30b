mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/n_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n_smoke.log 2>&1; echo "smoke rc=$?"; grep smoke gpurun_out/n_smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-extra-legs --no-cpu-baseline > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/n_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/n_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','p50_step_latency_ms','clocks','gpu_launches'): print(k, d.get(k))
PY
timeout 300 python tools/ill_step_cost.py pointmaze 4096 2>&1 | tail -4 | tee gpurun_out/n_ill.log
for mb in 4096 8192; do
  echo "== proj tc min batch $mb"
  DAD_TUNING=1 DAD_PROJ_TC_MIN_B=$mb DAD_LIB_PATH=$PWD/dynamics_aware_diffusion_b200/libdad_tune.so timeout 300 python tools/fusion_sweep.py pointmaze 4096 2>&1 | grep "B=  4096" | tee -a gpurun_out/n_projab.log
done
