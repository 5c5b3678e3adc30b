"""Benchmark of the reverse-diffusion sampling path (BASELINE.json: plans/sec, dynamics-aware, B=4096, H=32,
500 steps; p50 diffusion-step latency).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 3 --warmup 1     # the reference's CPU sampler (torch port) on host cores

One bench "step" = one full plan batch: B trajectories taken through all S reverse-diffusion steps (U-Net +
fused step/projection/inpaint kernel per step, CUDA-graph replays).  `value` = N*B*K / time, time measured with
CUDA events on the launching stream between barriers, max over ranks; inputs resident in HBM.  `e2e` = the same
through the C-ABI host-buffer call (dad_sample_host): pinned host x_S in, host x_0 out, every plan batch.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

# name -> reference configuration (SURVEY.md 8(d) "Configs -> concrete shapes")
WORKLOADS = {
    "pointmaze": dict(n=4, m=2, dim=128, mults=(1, 2, 4), H=32, S=500, B=4096, dyn="double_integrator",
                      label="PointMaze UMaze dynamics-aware (known double-integrator projector)"),
    "pointmaze_guided": dict(n=4, m=2, dim=128, mults=(1, 2, 4), H=32, S=100, B=64, dyn=None,
                             label="PointMaze UMaze guided sampling (config 0, the reference's CPU-runnable case)"),
    "halfcheetah": dict(n=17, m=6, dim=256, mults=(1, 4, 8), H=32, S=1000, B=1024, dyn="data_driven",
                        label="HalfCheetah dynamics-aware, 1024 plans per GPU"),
    "door": dict(n=39, m=28, dim=256, mults=(1, 2, 4, 8), H=32, S=1000, B=4096, dyn="data_driven",
                 label="AdroitHand Door data-driven projector"),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def ncu_traffic(kernel, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed `ncu --set full`
    summary of the same workload (profiles/ncu_traffic.json, written by tools/summarize_ncu.py); None if absent."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    try:
        return json.load(open(path)).get(workload, {}).get(kernel)
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                mask = get(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._halt.wait(0.1)

    def finish(self):
        self._halt.set()
        self.join(2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def projector_inputs(w):
    """(P fp32 torch, normalizer) for the workload, via the host-side builders (ProjectionMatrixBuilder,
    fit_linear_dynamics) on synthetic dynamics."""
    from dynamics_aware_diffusion_b200 import ProjectionMatrixBuilder, fit_linear_dynamics, synthetic
    nz = synthetic.SyntheticNormalizer(w["n"], w["m"], seed=7)
    if w["dyn"] is None:
        return None, nz
    if w["dyn"] == "double_integrator":
        A, B = synthetic.double_integrator(0.1)
    else:
        _, _, X, U, Xn = synthetic.random_linear_system(w["n"], w["m"], seed=11, n_transitions=100_000)
        A, B = fit_linear_dynamics(X, U, Xn)
    return ProjectionMatrixBuilder(A, B, w["n"], w["m"]).get_projection_matrix(w["H"]), nz


def stream_step_roofline(dev, pk, B=262144, H=32, T=6):
    """K7 (x0, clamp, posterior mean, noise, inpainting; no projector) at B=262144: 12*H*T*B = 604 MB per launch with
    in-kernel Philox, 805 MB with injected noise -- several times the 126 MB L2, so every byte comes from HBM."""
    import torch
    from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion, _native as N
    net = TemporalUnet(T, dim=64, dim_mults=(1,), precision="bf16", max_batch=B)      # small arena; the U-Net is not run
    dif = GaussianDiffusion(net, horizon=H, observation_dim=T - 2, action_dim=2, n_timesteps=100).to(dev)
    dif.fp32_ill_conditioned_steps = False      # only the step kernels run here: no fp32 sibling (and its arena) needed
    eng = dif.engine(H, dev)
    eng.set_conditions({0: torch.zeros(1, T, device=dev)}, B)
    out = {"bound": "hbm", "peak": pk["hbm"], "unit": "GB/s", "kernel": "step_pointwise_kernel",
           "workload": "B=%d H=%d T=%d, start inpainting" % (B, H, T)}
    # in-kernel Philox noise: 12 B/element (read x, eps; write x); injected noise (parity mode): 16 B/element
    for key, extra, per_elt in (("philox", 0, 12), ("injected_noise", 0x100, 16)):
        ms = eng.time_step_kernel(B, 50, flags=N.FLAG_CONDITIONS | extra, iters=20)
        nbytes = per_elt * H * T * B
        out[key] = {"achieved": nbytes / (ms * 1e-3) / 1e9, "frac": nbytes / (ms * 1e-3) / 1e9 / pk["hbm"],
                    "avg_launch_ms": ms, "bytes_per_launch": nbytes}
    out["achieved"], out["frac"] = out["injected_noise"]["achieved"], out["injected_noise"]["frac"]
    # The kernel the HEADLINE config runs -- step_project_fused_kernel, pointwise part + the D x D projector from shared
    # memory -- at the same streaming size.  2*D*D = 73.7 kFLOP per 2.3 KB sample: 32 FLOP/B, three times the fp32-SIMT
    # ridge of this machine, so at streaming sizes it is bound by the FMA pipe, not by HBM; reported as measured.
    if H * T == WORKLOADS["pointmaze"]["H"] * (WORKLOADS["pointmaze"]["n"] + WORKLOADS["pointmaze"]["m"]):
        pol, eng2, pflags, _, _ = attach_policy(dif, dict(WORKLOADS["pointmaze"], S=100), dev, B)
        D = H * T
        nbytes = 12 * D * B
        fp = {}
        # 0x400: the fused SIMT kernel (projector in shared memory); 0x200: pointwise kernel + bf16x3 tcgen05 GEMM (K8)
        for key, extra, kern in (("fused_simt", 0x400, "step_project_fused_kernel"),
                                 ("tensor_core", 0x200, "step_pointwise_kernel + conv_tc_kernel<128,0> (bf16x3 projector GEMM)")):
            ms = eng2.time_step_kernel(B, 50, flags=pflags | extra, iters=10)
            fp[key] = {"kernel": kern, "avg_launch_ms": ms, "achieved": nbytes / (ms * 1e-3) / 1e9,
                       "frac": nbytes / (ms * 1e-3) / 1e9 / pk["hbm"], "projector_tflops": 2.0 * D * D * B / (ms * 1e-3) / 1e12}
        fp["bytes_per_launch"] = nbytes
        fp["note"] = ("algorithmic bytes 12*D per sample (read x, eps; write x); the projector adds 2*D*D FLOP per sample = "
                      "%.0f FLOP/B: fp32-FMA-bound on SIMT, so large batches take the tensor-core path" % (2.0 * D * D / (12 * D)))
        out["fused_projector"] = fp
        del pol, eng2
    # DRAM bytes of one injected-noise launch at B=262144 (profiles/r01_ncu_full_stream_b262144.md)
    out["traffic"] = ncu_traffic("step_pointwise_kernel", "stream_b262144") if B == 262144 else None
    del eng, dif, net
    torch.cuda.empty_cache()
    return out


def workload_config(name, w, world, scaling):
    """The `config` object of the JSON line: identical in both arms (b200 and reference) by construction."""
    B_total = w["B"] if (scaling == "strong" or world == 1) else w["B"] * world
    return {"workload": name, "description": w["label"], "B_total": B_total, "B_per_gpu": B_total // world, "H": w["H"],
            "T": w["n"] + w["m"], "diffusion_steps": w["S"], "policy": "dynamics-aware" if w["dyn"] is not None else "guided",
            "unet": "dim=%d mults=%s" % (w["dim"], ",".join(map(str, w["mults"]))),
            "noise": "independent Gaussian per plan and step (GPU arm: in-kernel Philox; reference arm: torch.randn on the host)",
            "parallelism": "batch sharded over the GPUs, replicated weights (reference arm: the host cores of one box, rank 0 only)",
            "l2": "no flush: a diffusion step's working set (activations + weights, `working_set_mb` of the GPU arm's line) "
                  "is produced and consumed inside the step and exceeds the 126 MB L2 from B=1024 per GPU up"}


def run_reference(args, w, name):
    """The reference's CPU sampler on the host cores.  With oracle/_ref staged (oracle/make_ref.py: the reference's own
    files, byte for byte) the UNMODIFIED reference classes run -- `p_sample_with_guidance` -> `apply_projection` ->
    `apply_conditions` per step, the composition the goldens pin (tests/golden/make_golden.py) -- and the line says
    kind "reference"; otherwise the pinned restatement oracle/torch_port.py runs (kind "port")."""
    import contextlib
    import io
    import numpy as np
    import torch
    from oracle import torch_port, ref_shim
    from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion, synthetic, projection_alphas
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    T = w["n"] + w["m"]
    S, H = w["S"], w["H"]
    net = TemporalUnet(T, dim=w["dim"], dim_mults=w["mults"])
    dif = GaussianDiffusion(net, horizon=H, observation_dim=w["n"], action_dim=w["m"], n_timesteps=S)
    synthetic.fill_state_dict(dif, 0)
    sd = {k: v.detach() for k, v in dif.state_dict().items()}
    P, nz = projector_inputs(w)
    Bc, dsteps = args.cpu_batch, args.cpu_diffusion_steps
    g = torch.Generator().manual_seed(1234)
    start = torch.zeros(T)
    start[:w["n"]] = torch.randn(w["n"], generator=g)
    kind = "reference" if ref_shim.available() and not args.cpu_port else "port"

    if kind == "reference":
        ref = ref_shim.load()
        rnet = ref.TemporalUnet(T, dim=w["dim"], dim_mults=w["mults"])
        rdif = ref.GaussianDiffusion(rnet, horizon=H, observation_dim=w["n"], action_dim=w["m"], n_timesteps=S)
        rdif.load_state_dict(sd, strict=True)
        rdif.eval()
        with contextlib.redirect_stdout(io.StringIO()):
            if P is not None:
                rpol = ref.DynamicsAwarePolicy(rdif, projection_matrix=P, normalizer=nz, state_dim=w["n"],
                                               observation_dim=w["n"], action_dim=w["m"], horizon=H,
                                               projection_schedule="noise_schedule", projection_strength=1.0)
            else:
                rpol = ref.GuidedPolicy(rdif, nz)
        cond = {0: start[None]}

        @torch.no_grad()
        def one_sample():
            # torch.randn inside p_sample_with_guidance draws the step noise, as in the reference
            x = rpol.apply_conditions(torch.randn(Bc, H, T, generator=g), cond)
            for i in reversed(range(S - dsteps, S)):
                t = torch.full((Bc,), i, dtype=torch.long)
                if P is not None:          # denoise -> project -> inpaint (README.md:24-25)
                    x = rpol.p_sample_with_guidance(x, t, None)
                    x = rpol.apply_projection(x, i)
                    x = rpol.apply_conditions(x, cond)
                else:
                    x = rpol.p_sample_with_guidance(x, t, cond)
            return x
    else:
        projector = None
        if P is not None:
            al = projection_alphas(S, S, "noise_schedule", 1.0, sd["betas"])
            projector = dict(P=P, alphas=[float(a) for a in al], n=w["n"], m=w["m"], H=H,
                             nz=tuple(torch.from_numpy(np.asarray(a, dtype=np.float32))
                                      for a in (nz.obs_mean, nz.obs_std, nz.action_mean, nz.action_std)))

        def one_sample():
            x = torch.randn(Bc, H, T, generator=g)
            return torch_port.sample_loop(sd, x, lambda k: torch.randn(Bc, H, T, generator=g), {0: start}, projector,
                                          steps=dsteps)

    for _ in range(args.warmup):
        one_sample()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_sample()
    dt = time.perf_counter() - t0
    per_dstep = dt / (args.steps * dsteps)
    value = Bc / (per_dstep * S)
    sample = "B=%d plans x %d of %d diffusion steps per bench step, extrapolated linearly to %d steps" % (Bc, dsteps, S, S)
    line = {
        "impl": "reference", "metric": "plans/sec", "value": value, "unit": "plans/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak" if (args.gpus > 1 and args.scaling == "weak") else "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(name, w, args.gpus, args.scaling),
        "cpu_baseline": {"value": value, "unit": "plans/s", "cores": cores, "kind": kind, "sample": sample,
                         "code": ("the reference's own modules (oracle/_ref, staged by oracle/make_ref.py)" if kind == "reference"
                                  else "oracle/torch_port.py (restatement pinned to the reference by tests/golden)"),
                         "ms_per_diffusion_step_at_sample_B": per_dstep * 1e3},
        "e2e": {"value": value, "unit": "plans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def build_policy(w, B, precision, dev, rank=0, latency_max_batch=None):
    """(policy, diffusion, engine, flags, start condition) for a workload with random-init weights (seed 0 on rank 0)."""
    import torch
    from dynamics_aware_diffusion_b200 import (TemporalUnet, GaussianDiffusion, GuidedPolicy, DynamicsAwarePolicy,
                                               synthetic, _native as N)
    T, S, H = w["n"] + w["m"], w["S"], w["H"]
    net = TemporalUnet(T, dim=w["dim"], dim_mults=w["mults"], precision=precision, max_batch=min(B, w.get("max_batch", B)),
                       latency_max_batch=latency_max_batch)
    dif = GaussianDiffusion(net, horizon=H, observation_dim=w["n"], action_dim=w["m"], n_timesteps=S)
    synthetic.fill_state_dict(dif, 0 if rank == 0 else 1000 + rank)
    dif.to(dev)
    return net, dif


def attach_policy(dif, w, dev, B):
    import torch
    from dynamics_aware_diffusion_b200 import GuidedPolicy, DynamicsAwarePolicy, _native as N
    T, H = w["n"] + w["m"], w["H"]
    P, nz = projector_inputs(w)
    if P is not None:
        pol = DynamicsAwarePolicy(dif, projection_matrix=P, normalizer=nz, state_dim=w["n"], observation_dim=w["n"],
                                  action_dim=w["m"], horizon=H, projection_schedule="noise_schedule",
                                  projection_strength=1.0)
    else:
        pol = GuidedPolicy(dif, nz)
    eng = pol._engine(dev)
    flags = pol._loop_flags(eng) | N.FLAG_CONDITIONS
    start = torch.zeros(1, T, device=dev)
    start[0, :w["n"]] = torch.randn(w["n"], generator=torch.Generator().manual_seed(1234)).to(dev)
    eng.set_conditions({0: start}, B)
    return pol, eng, flags, start, P is not None


def config_leg(name, dev, pk, overrides, dsteps=10):
    """One of BASELINE.json's other configurations, measured briefly on one GPU: `dsteps` diffusion steps of the
    configured sampler at the configured batch (graph replays timed individually with CUDA events), reported as p50
    step latency, plans/s extrapolated to the configured number of steps, and the U-Net tensor fraction."""
    import torch
    from dynamics_aware_diffusion_b200 import _native as N
    w = dict(WORKLOADS[name])
    w.update(overrides)
    B, S = w["B"], w["S"]
    cap = min(B, w.get("max_batch", B))
    t0 = time.perf_counter()
    net, dif = build_policy(w, cap, "bf16", dev)
    pol, eng, flags, start, dyn = attach_policy(dif, w, dev, cap)
    x = torch.empty(cap, w["H"], w["n"] + w["m"], device=dev)
    # the schedule has S entries; time the LAST dsteps indices (well-conditioned steps; every step runs the same kernels)
    eng.sample_profile(x, 3, flags=flags | N.FLAG_PHILOX_INIT, seed=1)           # warm-up + graph capture
    step_ms = eng.sample_profile(x, dsteps, flags=flags | N.FLAG_PHILOX_INIT, seed=2)
    p50 = statistics.median(step_ms)
    info = eng.info()
    chunks = (B + cap - 1) // cap
    out = {"workload": name, "B": B, "H": w["H"], "T": w["n"] + w["m"], "diffusion_steps": S,
           "unet": "dim=%d mults=%s" % (w["dim"], ",".join(map(str, w["mults"]))),
           "policy": "dynamics-aware" if dyn else "guided", "timed_diffusion_steps": dsteps,
           "p50_step_latency_ms": p50 * chunks, "plans_per_s_extrapolated": B / (p50 * chunks * 1e-3 * S),
           "unet_tensor_frac_of_sustained": info["conv_flops_per_sample"] * cap / (p50 * 1e-3) / (pk["tf_sustained"] * 1e12),
           "launches_per_step": info["launches_per_step"], "setup_s": None}
    assert bool(torch.isfinite(x).all()), "non-finite trajectories in leg %s" % name
    if chunks > 1:
        out["note"] = "B exceeds the %d-plan workspace: %d chunks per step, latency = chunks x measured" % (cap, chunks)
    del eng, pol, dif, net, x
    torch.cuda.empty_cache()
    out["setup_s"] = round(time.perf_counter() - t0, 1)
    return out


def sharded_leg(name, dev, pk, world, dist, B_total, dsteps=10):
    """BASELINE.json configs 3 / 4 on N GPUs (every rank calls this): the batch B_total split evenly over the ranks, each
    rank times `dsteps` diffusion steps of its shard (config_leg), the step latency is the MAX over ranks, and the final
    trajectory all-gather (C2) is timed on its own and added once per plan."""
    import torch
    w = WORKLOADS[name]
    Bg = B_total // world
    out = config_leg(name, dev, pk, {"B": Bg, "max_batch": Bg}, dsteps)
    t = torch.tensor([out["p50_step_latency_ms"]], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    p50 = float(t.item())
    T = w["n"] + w["m"]
    x = torch.zeros(Bg, w["H"], T, device=dev)
    g = torch.empty(Bg * world, w["H"], T, device=dev)
    dist.all_gather_into_tensor(g, x)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        dist.all_gather_into_tensor(g, x)
    e1.record()
    torch.cuda.synchronize()
    gt = torch.tensor([e0.elapsed_time(e1) / 3], device=dev, dtype=torch.float64)
    dist.all_reduce(gt, op=dist.ReduceOp.MAX)
    S = w["S"]
    plan_ms = p50 * S + float(gt.item())
    return {"workload": name, "n_gpus": world, "B_total": Bg * world, "B_per_gpu": Bg, "H": w["H"], "T": T,
            "diffusion_steps": S, "unet": out["unet"], "policy": out["policy"], "timed_diffusion_steps": dsteps,
            "p50_step_latency_ms_max_over_ranks": p50, "gather_ms": float(gt.item()),
            "plans_per_s_extrapolated": Bg * world / (plan_ms * 1e-3),
            "unet_tensor_frac_of_sustained_per_gpu": out["unet_tensor_frac_of_sustained"] * out["p50_step_latency_ms"] / p50}


def config1_full_leg(dev, with_cpu=True):
    """BASELINE.json configs[0] -- PointMaze guided sampling, B=64, H=32, 100 steps: the reference's own CPU-runnable case --
    run IN FULL on both sides, nothing extrapolated: `GuidedPolicy.sample_loop(batch_size=64, conditions)` of this package on
    the GPU (median of 5 loops after 2 warm-up loops, wall clock around the call incl. the final synchronize) and the
    reference's own `GuidedPolicy.sample_loop` (oracle/_ref, else the torch port) on the host cores (1 loop)."""
    import contextlib
    import io
    import torch
    from oracle import ref_shim, torch_port
    from dynamics_aware_diffusion_b200 import GuidedPolicy
    w = WORKLOADS["pointmaze_guided"]
    B, S, H, T = w["B"], w["S"], w["H"], w["n"] + w["m"]
    net, dif = build_policy(w, B, "bf16", dev)
    _, nz = projector_inputs(w)
    pol = GuidedPolicy(dif, nz)
    start = torch.zeros(1, T, device=dev)
    start[0, :w["n"]] = torch.randn(w["n"], generator=torch.Generator().manual_seed(1234)).to(dev)
    lat = []
    for k in range(7):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x = pol.sample_loop(batch_size=B, conditions={0: start}, seed=k)
        torch.cuda.synchronize()
        lat.append(time.perf_counter() - t0)
    assert bool(torch.isfinite(x).all()) and bool((x[:, 0] == start).all())
    gpu_s = statistics.median(lat[2:])
    out = {"workload": "pointmaze_guided", "B": B, "H": H, "T": T, "diffusion_steps": S, "extrapolated": False,
           "gpu": {"loop_ms": gpu_s * 1e3, "plans_per_s": B / gpu_s, "ms_per_diffusion_step": gpu_s * 1e3 / S}}
    if with_cpu:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sd = {k: v.detach().cpu() for k, v in dif.state_dict().items()}
        g = torch.Generator().manual_seed(7)
        kind = "reference" if ref_shim.available() else "port"
        if kind == "reference":
            ref = ref_shim.load()
            rnet = ref.TemporalUnet(T, dim=w["dim"], dim_mults=w["mults"])
            rdif = ref.GaussianDiffusion(rnet, horizon=H, observation_dim=w["n"], action_dim=w["m"], n_timesteps=S)
            rdif.load_state_dict(sd, strict=True)
            rdif.eval()
            with contextlib.redirect_stdout(io.StringIO()):
                rpol = ref.GuidedPolicy(rdif, nz)
            fn = lambda: rpol.sample_loop(batch_size=B, conditions={0: start.cpu()})
        else:
            fn = lambda: torch_port.sample_loop(sd, torch.randn(B, H, T, generator=g),
                                                lambda k: torch.randn(B, H, T, generator=g), {0: start.cpu()[0]}, None)
        with torch.no_grad():
            t0 = time.perf_counter()
            y = fn()
            cpu_s = time.perf_counter() - t0
        assert bool(torch.isfinite(torch.as_tensor(y)).all())
        out["cpu_reference"] = {"loop_ms": cpu_s * 1e3, "plans_per_s": B / cpu_s, "cores": cores, "kind": kind}
        out["gpu_over_cpu_reference"] = cpu_s / gpu_s
    del pol, dif, net
    torch.cuda.empty_cache()
    return out


def gpu_eager_baseline(w, dev, B, dsteps=3):
    """The honest GPU comparator: the reference's op sequence in STOCK PyTorch eager
    (oracle/torch_port.py: cuDNN convs, native GroupNorm / Mish, the 15-op projection chain, torch.randn noise) on the
    same B200, same batch, fp32 (PyTorch defaults: TF32 convs) and bf16 autocast.  `dsteps` steps timed with CUDA
    events after one warm-up pass, extrapolated to the configured number of diffusion steps."""
    import numpy as np
    import torch
    from oracle import torch_port
    from dynamics_aware_diffusion_b200 import TemporalUnet, GaussianDiffusion, synthetic, projection_alphas
    T, S, H = w["n"] + w["m"], w["S"], w["H"]
    net = TemporalUnet(T, dim=w["dim"], dim_mults=w["mults"])
    dif = GaussianDiffusion(net, horizon=H, observation_dim=w["n"], action_dim=w["m"], n_timesteps=S)
    synthetic.fill_state_dict(dif, 0)
    sd = {k: v.detach().to(dev) for k, v in dif.state_dict().items()}
    P, nz = projector_inputs(w)
    projector = None
    if P is not None:
        al = projection_alphas(S, S, "noise_schedule", 1.0, sd["betas"].cpu())
        projector = dict(P=P.to(dev), alphas=[float(a) for a in al], n=w["n"], m=w["m"], H=H,
                         nz=tuple(torch.from_numpy(np.asarray(a, dtype=np.float32)).to(dev)
                                  for a in (nz.obs_mean, nz.obs_std, nz.action_mean, nz.action_std)))
    start = torch.zeros(T, device=dev)
    out = {"what": "torch eager (oracle/torch_port.py) on this GPU, B=%d, %d diffusion steps timed, extrapolated to %d"
                   % (B, dsteps, S), "cudnn_allow_tf32": bool(torch.backends.cudnn.allow_tf32)}
    for key, ctx in (("fp32", None), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
        def run(steps):
            x = torch.randn(B, H, T, device=dev)
            fn = lambda: torch_port.sample_loop(sd, x, lambda k: torch.randn(B, H, T, device=dev), {0: start}, projector,
                                                steps=steps)
            if ctx is None:
                return fn()
            with ctx:
                return fn()
        run(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(dsteps)
        e1.record()
        torch.cuda.synchronize()
        ms_step = e0.elapsed_time(e1) / dsteps
        out[key] = {"ms_per_diffusion_step": ms_step, "plans_per_s": B / (ms_step * 1e-3 * S)}
    del sd
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="pointmaze", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="TOTAL plans per bench step (default: the workload's)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = the workload's B split over the ranks (BASELINE.json's metric: B=4096 at 1/2/4/8 "
                         "GPUs), weak = the workload's B on every rank")
    ap.add_argument("--diffusion-steps", type=int, default=0, help="override S (default: the workload's)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-batch", type=int, default=256)
    ap.add_argument("--cpu-diffusion-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-port", action="store_true", help="reference arm: time oracle/torch_port.py even when oracle/_ref is staged")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the other BASELINE configs, the sweep and the eager-PyTorch comparator")
    ap.add_argument("--layers-out", default="", help="write the per-layer / per-launch timing tables (JSON) to this file")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.batch:
        w["B"] = args.batch
    if args.diffusion_steps:
        w["S"] = args.diffusion_steps
    if args.impl == "reference":
        return run_reference(args, w, args.workload)
    if args.warmup < 3:
        print("note: fewer than 3 warm-up steps requested; the timing rules ask for >= 3", file=sys.stderr)

    import numpy as np
    import torch
    import torch.distributed as dist
    from dynamics_aware_diffusion_b200 import _native as N
    from dynamics_aware_diffusion_b200.distributed import broadcast_module, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU: this implementation has no CPU path"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    T, S, H = w["n"] + w["m"], w["S"], w["H"]
    B_total = w["B"] if (args.scaling == "strong" or world == 1) else w["B"] * world
    lo, hi = shard_bounds(B_total, rank, world)
    B = hi - lo                                   # this rank's plans
    B_max = shard_bounds(B_total, 0, world)[1]    # the largest shard (all_gather_into_tensor wants equal shards)
    assert B_total % world == 0, "the batch must divide over the ranks (all_gather_into_tensor)"

    # ---- model: random-init weights of the configured architecture, rank 0's replicated over NCCL (C1)
    net, dif = build_policy(w, w["B"], args.precision, dev, rank, latency_max_batch=0 if world > 1 else None)
    if world > 1:
        torch.cuda.nvtx.range_push("broadcast_weights")
        broadcast_module(dif, src=0)
        torch.cuda.nvtx.range_pop()
    pol, eng, flags, start, dyn = attach_policy(dif, w, dev, B)
    info = eng.info()
    x = torch.empty(B, H, T, device=dev)
    gathered = torch.empty(B_total, H, T, device=dev) if world > 1 else None

    def plan_batch(k, xb=None, out=None):
        # x_S drawn in-kernel (Philox, subsequence = global sample index), then S graph replays
        xb = x if xb is None else xb
        torch.cuda.nvtx.range_push("sample_loop")
        eng.sample(xb, S, flags=flags | N.FLAG_PHILOX_INIT, seed=1234 + k, sample_offset=lo)
        torch.cuda.nvtx.range_pop()
        if world > 1:
            torch.cuda.nvtx.range_push("gather_trajectories")
            dist.all_gather_into_tensor(gathered if out is None else out, xb)   # C2: final trajectory gather over NVLink
            torch.cuda.nvtx.range_pop()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(n_steps):
            fn(100 + k)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for k in range(args.warmup):
        plan_batch(k)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = eng.launch_count()
    ms_total = timed(args.steps, plan_batch)
    clocks = sampler.finish()
    launches = eng.launch_count() - launches0
    value = B_total * args.steps / (ms_total * 1e-3)
    assert bool(torch.isfinite(x).all()), "sampler produced non-finite trajectories"
    assert bool((x[:, 0] == start).all()), "inpainting lost"

    # ---- multi-GPU correctness: the gathered batch equals a single-GPU recomputation, bit for bit.  The Philox
    # subsequence of a plan is its GLOBAL index and a plan's result does not depend on its batch, so rank 0 recomputes
    # 8 plans of EVERY rank's shard alone (sample_offset = their global index) and compares with what NCCL delivered.
    nccl_check = None
    if world > 1:
        plan_batch(999)
        torch.cuda.synchronize()
        assert torch.equal(gathered[lo:hi], x), "all_gather misplaced this rank's shard"
        if rank == 0:
            nccl_check = {"rows_per_rank": 8, "ranks": world, "bit_exact": True}
            for r in range(world):
                r0 = shard_bounds(B_total, r, world)[0]
                xr = torch.empty(8, H, T, device=dev)
                eng.set_conditions({0: start}, 8)
                eng.sample(xr, S, flags=flags | N.FLAG_PHILOX_INIT, seed=1234 + 999, sample_offset=r0)
                torch.cuda.synchronize()
                if not torch.equal(xr, gathered[r0:r0 + 8]):
                    nccl_check["bit_exact"] = False
            eng.set_conditions({0: start}, B)
            assert nccl_check["bit_exact"], "gathered trajectories differ from the single-GPU recomputation"
        dist.barrier()

    # ---- the other scaling mode as an extra (N > 1): weak = the workload's B on every rank
    other = None
    if world > 1 and B_total // world != w["B"]:
        Bw = w["B"]
        xw = torch.empty(Bw, H, T, device=dev)
        gw = torch.empty(world * Bw, H, T, device=dev)
        eng.set_conditions({0: start}, Bw)
        lo_keep = lo
        lo = rank * Bw
        plan_batch(0, xw, gw)
        ms_w = timed(2, lambda k: plan_batch(k, xw, gw))
        lo = lo_keep
        eng.set_conditions({0: start}, B)
        other = {"scaling": "weak", "B_per_gpu": Bw, "value": world * Bw * 2 / (ms_w * 1e-3), "unit": "plans/s", "steps": 2}
        del xw, gw

    # ---- e2e: the C-ABI host-buffer call; pinned x_S in, x_0 out, every plan batch
    xh = torch.randn(B, H, T, generator=torch.Generator().manual_seed(7 + rank)).pin_memory()
    xs = xh.clone().pin_memory()
    eng.sample_host(xs, S, flags=flags, seed=99, sample_offset=lo)        # warm-up (allocations)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        xs.copy_(xh)
        eng.sample_host(xs, S, flags=flags, seed=200 + k, sample_offset=lo)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = B_total * args.steps / float(e2e_s.item())
    bytes_io = B * H * T * 4

    line = {
        "metric": "plans/sec", "value": value, "unit": "plans/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak" if (world > 1 and args.scaling == "weak") else "strong",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": workload_config(args.workload, w, world, args.scaling),
        "working_set_mb": round(info["workspace_bytes"] / 1e6 * B / max(w["B"], 1), 1),
        "e2e": {"value": e2e_value, "unit": "plans/s", "h2d_bytes_per_step": bytes_io, "d2h_bytes_per_step": bytes_io},
        "gpu_launches": int(launches), "launches_per_diffusion_step": info["launches_per_step"], "clocks": clocks,
    }
    if nccl_check:
        line["nccl_check"] = nccl_check
    if other:
        line["other_scaling"] = other

    # ---- BASELINE.json configs 3 and 4 on N GPUs: HalfCheetah B = 1024 per GPU (8192 over 8), Door B = 4096 split N ways
    if world > 1 and not args.no_extra_legs:
        pk_all = peaks()
        multi = []
        for name, bt in (("halfcheetah", WORKLOADS["halfcheetah"]["B"] * world), ("door", WORKLOADS["door"]["B"])):
            try:
                multi.append(sharded_leg(name, dev, pk_all, world, dist, bt))
            except Exception as exc:      # a failing extra must not take the headline line with it (all ranks fail alike)
                multi.append({"workload": name, "error": str(exc)[:300]})
        line["configs_multi_gpu"] = multi

    if rank == 0:
        pk = peaks()
        # ---- p50 diffusion-step latency (CUDA events between graph replays)
        step_ms = eng.sample_profile(x, S, flags=flags | N.FLAG_PHILOX_INIT, seed=5)
        line["p50_step_latency_ms"] = statistics.median(step_ms)
        line["p99_step_latency_ms"] = sorted(step_ms)[int(0.99 * (len(step_ms) - 1))]
        flops_step = info["conv_flops_per_sample"] * B
        per_step_s = ms_total * 1e-3 / (args.steps * S)
        # the whole step is timed inside the long timed region: the SUSTAINED peak is its denominator
        line["unet_tensor_frac_of_sustained"] = flops_step / per_step_s / (pk["tf_sustained"] * 1e12)
        # ---- per-launch timing: the dominant kernel = the launch unit (a conv chain or a single conv) with the
        # largest share of the U-Net time.  Units are timed ALONE (back-to-back launches of one unit): the BURST peak
        # is the denominator for those.
        units = eng.units()
        for u in units:
            u["ms"] = eng.time_unit(u["index"], B, iters=20)
            u["tflops"] = u["flops_per_sample"] * B / (u["ms"] * 1e-3) / 1e12
        unet_ms = sum(u["ms"] for u in units)
        dom = max(units, key=lambda u: u["ms"])
        achieved = dom["tflops"]
        layers = eng.layers()
        shapes = ", ".join(sorted({"L=%d K=%d N=%d" % (l["L_out"], l["C_in"] * l["taps"], l["C_out"])
                                   for l in layers[dom["first_layer"]:dom["first_layer"] + dom["n_layers"]]}))
        line["roofline"] = {
            "bound": "tensor", "achieved": achieved, "peak": pk["tf_burst"], "unit": "TFLOP/s",
            "frac": achieved / pk["tf_burst"], "frac_of_sustained": achieved / pk["tf_sustained"],
            "traffic": ncu_traffic(dom["kernel"].split("<")[0], args.workload),
            "peak_source": pk["source"] + " burst bf16 (the kernel is timed alone, back to back)",
            "kernel": "%s: %d convs in one launch (%s)" % (dom["kernel"], dom["n_layers"], shapes),
            "share_of_unet_time": dom["ms"] / unet_ms, "flops_per_launch": dom["flops_per_sample"] * B,
            "avg_launch_ms": dom["ms"],
        }
        # ---- the fused step kernel against the HBM roofline
        st_ms = eng.time_step_kernel(B, S // 2, flags=flags, iters=50)
        st_bytes = 12 * H * T * B          # read x, read eps, write x (Philox noise): SURVEY.md 8(d)
        line["roofline_step_kernel"] = {
            "bound": "hbm", "achieved": st_bytes / (st_ms * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
            "frac": st_bytes / (st_ms * 1e-3) / 1e9 / pk["hbm"],
            "traffic": ncu_traffic("step_project_fused_kernel", args.workload), "avg_launch_ms": st_ms,
            "bytes_per_launch": st_bytes, "note": "working set %.1f MB is L2-resident at this B" % (st_bytes / 1e6)}
        # the same memory-bound step kernel without a projector (guided / plain policy) on a batch whose working set
        # leaves the L2 (the HBM fraction is only observable there)
        try:
            line["roofline_step_kernel_stream"] = stream_step_roofline(dev, pk)
        except Exception as exc:      # measurement extra: never fail the bench line over it
            line["roofline_step_kernel_stream"] = {"error": str(exc)[:200]}
        line["unet_ms_sum_of_launches"] = unet_ms
        # the production caller's shape (GuidedPolicy.get_action, policies.py:193-223): ONE plan, latency-bound
        if world == 1:
            try:
                def one_plan_ms():
                    lat = []
                    for k in range(4):
                        torch.cuda.synchronize()
                        t0 = time.perf_counter()
                        one = pol.sample_loop(batch_size=1, conditions={0: start}, seed=k)
                        _ = one[0, :2].cpu()                      # the actions get_action reads back
                        lat.append((time.perf_counter() - t0) * 1e3)
                    return statistics.median(lat[1:])
                p50 = one_plan_ms()                               # default: the latency kernels (conv_small) up to B = 24
                eng.set_latency_batch(0)
                p50_tp = one_plan_ms()                            # the same plan through the throughput kernels
                eng.set_latency_batch(24)
                line["plan_latency_b1_ms"] = {"p50": p50, "p50_throughput_kernels": p50_tp, "diffusion_steps": S,
                                              "us_per_diffusion_step": p50 * 1e3 / S,
                                              "note": "sample_loop(batch_size=1) + D2H of the first actions, wall clock"}
                eng.set_conditions({0: start}, B)
            except Exception as exc:
                line["plan_latency_b1_ms"] = {"error": str(exc)[:200]}
        if args.layers_out:
            os.makedirs(os.path.dirname(os.path.abspath(args.layers_out)), exist_ok=True)
            json.dump({"units": units, "layers": layers}, open(args.layers_out, "w"), indent=1)
        if world == 1 and not args.no_extra_legs:
            # free the headline model first: the other configurations need their own workspaces
            del eng, pol
            net._engines.clear()
            torch.cuda.empty_cache()
            # ---- the honest GPU comparator: stock PyTorch eager on this B200
            try:
                line["gpu_eager_baseline"] = gpu_eager_baseline(w, dev, B)
                line["gpu_eager_baseline"]["speedup_vs_fp32"] = value / line["gpu_eager_baseline"]["fp32"]["plans_per_s"]
                line["gpu_eager_baseline"]["speedup_vs_bf16_autocast"] = value / line["gpu_eager_baseline"]["bf16_autocast"]["plans_per_s"]
            except Exception as exc:
                line["gpu_eager_baseline"] = {"error": str(exc)[:300]}
            # ---- BASELINE.json's other configurations and the scaling sweep, short legs on this GPU
            legs = [("pointmaze_guided", {}), ("halfcheetah", {}), ("door", {}),
                    ("pointmaze", {"H": 64}), ("pointmaze", {"H": 128, "B": 2048}),
                    ("pointmaze", {"B": 256}), ("pointmaze", {"B": 1024}), ("pointmaze", {"B": 16384, "max_batch": 16384}),
                    ("pointmaze", {"B": 65536, "max_batch": 16384})]
            try:
                line["config1_full"] = config1_full_leg(dev, with_cpu=not args.no_cpu_baseline)
            except Exception as exc:
                line["config1_full"] = {"error": str(exc)[:300]}
            line["configs"] = []
            for name, ov in legs:
                try:
                    line["configs"].append(config_leg(name, dev, pk, ov))
                except Exception as exc:
                    line["configs"].append({"workload": name, "overrides": ov, "error": str(exc)[:300]})
        # ---- CPU baseline: the reference's op sequence in torch on the host cores, bounded sample
        if not args.no_cpu_baseline and world == 1:
            import io
            import contextlib
            buf = io.StringIO()
            sub = argparse.Namespace(**vars(args))
            sub.steps, sub.warmup = 2, 1
            with contextlib.redirect_stdout(buf):
                run_reference(sub, w, args.workload)
            line["cpu_baseline"] = json.loads(buf.getvalue().strip().splitlines()[-1])["cpu_baseline"]
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
