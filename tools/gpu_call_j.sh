mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/j_pytest.log
timeout 600 python bench.py > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; echo "bench rc=$?"
timeout 200 python tools/plan_latency.py > gpurun_out/j_plan_latency.log 2>&1
bash tools/profile_round2.sh > gpurun_out/j_profile.log 2>&1; echo "profile rc=$?"
