mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/k_pytest.log
timeout 900 python bench.py > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/k_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/k_ref.json 2> gpurun_out/k_ref.err; echo "ref rc=$?"
