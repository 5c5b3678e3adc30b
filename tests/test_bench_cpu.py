"""The CPU-runnable legs of bench.py: the `--impl reference` arm (the reference's CPU sampler, torch port) prints the
contract's JSON line; workload table sanity."""
import json
import os
import subprocess
import sys

import helpers


def test_reference_arm_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(helpers.ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--cpu-batch", "8", "--cpu-diffusion-steps", "1"],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "plans/sec" and line["unit"] == "plans/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["config"]["workload"] == "pointmaze" and line["config"]["diffusion_steps"] == 500
    assert line["config"]["B_per_gpu"] == 4096 and line["config"]["policy"] == "dynamics-aware"
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert line["e2e"] == {"value": line["value"], "unit": "plans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_workload_table_matches_baseline_configs():
    sys.path.insert(0, helpers.ROOT)
    import bench
    w = bench.WORKLOADS
    assert (w["pointmaze"]["B"], w["pointmaze"]["H"], w["pointmaze"]["S"]) == (4096, 32, 500)       # BASELINE.json configs[1]
    assert w["pointmaze"]["dim"] == 128 and tuple(w["pointmaze"]["mults"]) == (1, 2, 4)
    assert tuple(w["halfcheetah"]["mults"]) == (1, 4, 8) and w["halfcheetah"]["n"] + w["halfcheetah"]["m"] == 23
    assert tuple(w["door"]["mults"]) == (1, 2, 4, 8) and w["door"]["n"] + w["door"]["m"] == 67
