"""The CPU-runnable legs of bench.py: the `--impl reference` arm (the reference's CPU sampler, torch port) prints the
contract's JSON line; workload table sanity."""
import json
import os
import subprocess
import sys

import helpers


import pytest


@pytest.mark.parametrize("port", [False, True])
def test_reference_arm_json_line(port):
    """Both flavours of the reference arm: the reference's own staged modules (oracle/_ref, when present) and the port."""
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(helpers.ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--cpu-batch", "8", "--cpu-diffusion-steps", "1"] + (["--cpu-port"] if port else []),
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "plans/sec" and line["unit"] == "plans/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["config"]["workload"] == "pointmaze" and line["config"]["diffusion_steps"] == 500
    assert line["config"]["B_per_gpu"] == 4096 and line["config"]["policy"] == "dynamics-aware"
    cb = line["cpu_baseline"]
    sys.path.insert(0, helpers.ROOT)
    from oracle import ref_shim
    assert cb["kind"] == ("reference" if ref_shim.available() and not port else "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert line["e2e"] == {"value": line["value"], "unit": "plans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_workload_table_matches_baseline_configs():
    sys.path.insert(0, helpers.ROOT)
    import bench
    w = bench.WORKLOADS
    assert (w["pointmaze"]["B"], w["pointmaze"]["H"], w["pointmaze"]["S"]) == (4096, 32, 500)       # BASELINE.json configs[1]
    assert w["pointmaze"]["dim"] == 128 and tuple(w["pointmaze"]["mults"]) == (1, 2, 4)
    assert tuple(w["halfcheetah"]["mults"]) == (1, 4, 8) and w["halfcheetah"]["n"] + w["halfcheetah"]["m"] == 23
    assert tuple(w["door"]["mults"]) == (1, 2, 4, 8) and w["door"]["n"] + w["door"]["m"] == 67


def test_staged_reference_is_the_reference_and_reproduces_a_golden_trace(tmp_path):
    """oracle/make_ref.py stages the path's files byte for byte; the classes imported from the STAGED copy reproduce the
    committed golden trace of the tiny case exactly (the goldens were made by the same code from /root/reference)."""
    import hashlib
    import numpy as np
    import torch
    sys.path.insert(0, helpers.ROOT)
    from oracle import make_ref, ref_shim
    staged = os.path.join(helpers.ROOT, "oracle", "_ref")
    if os.path.isdir(os.path.join(make_ref.SRC_ROOT, "m_diffuser", "models")):
        assert make_ref.stage(verbose=False)
        for rel in make_ref.FILES:
            a = hashlib.sha256(open(os.path.join(make_ref.SRC_ROOT, rel), "rb").read()).hexdigest()
            b = hashlib.sha256(open(os.path.join(staged, rel), "rb").read()).hexdigest()
            assert a == b, rel
    if not os.path.isdir(os.path.join(staged, "m_diffuser", "models")):
        pytest.skip("oracle/_ref not staged (no /root/reference here)")
    code = (
        "import sys, numpy as np, torch; sys.path[:0] = [%r, %r]\n"
        "import helpers\nfrom oracle import ref_shim\n"
        "assert ref_shim.REF_ROOT == %r, ref_shim.REF_ROOT\n"
        "ref = ref_shim.load(); c = helpers.CASES['tiny']; g = helpers.load_golden('tiny')\n"
        "sd, _ = helpers.make_state_dict(c); T = helpers.case_T(c)\n"
        "net = ref.TemporalUnet(T, dim=c['dim'], dim_mults=c['mults'])\n"
        "dif = ref.GaussianDiffusion(net, horizon=c['H'], observation_dim=c['n'], action_dim=c['m'], n_timesteps=c['S'], beta_schedule=c['beta'])\n"
        "dif.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); dif.eval()\n"
        "with torch.no_grad():\n"
        "    eps = net(torch.from_numpy(g['x_init']), torch.full((c['B'],), int(g['unet_steps'][0]), dtype=torch.long)).numpy()\n"
        "assert np.abs(eps - g['unet_eps'][0]).max() <= 1e-6 * np.abs(g['unet_eps'][0]).max(), 'staged reference differs from golden'\n"
        % (helpers.ROOT, os.path.join(helpers.ROOT, "tests"), staged))
    env = dict(os.environ, DAD_REFERENCE_ROOT=staged, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
