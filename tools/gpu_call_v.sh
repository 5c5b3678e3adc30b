mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/v_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/v_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v_smoke.log 2>&1; echo "smoke rc=$?"; grep smoke gpurun_out/v_smoke.log
timeout 300 python bench.py --impl reference > gpurun_out/v_ref.json 2> gpurun_out/v_ref.err; echo "ref rc=$?"
( time timeout 900 python bench.py > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -3 gpurun_out/v_bench.err
bash tools/profile_round2.sh > gpurun_out/v_profile.log 2>&1; echo "profile rc=$?"
