"""Handle-owning wrapper around the C ABI.  PyTorch is used for device memory and streams only."""
import ctypes
import warnings

import torch

from . import _native as N


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t, name):
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: this implementation has no CPU path" % name)
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32 (got %s)" % (name, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


class Engine:
    """One native sampler handle = (architecture, horizon, precision, batch capacity) on one device."""

    def __init__(self, *, transition_dim, dim, dim_mults, kernel_size, time_dim, horizon, n_timesteps,
                 precision, max_batch, device, predict_epsilon=True, clip_denoised=True):
        self.lib = N.lib()
        self.handle = ctypes.c_void_p()
        self.horizon, self.transition_dim = int(horizon), int(transition_dim)
        self.n_timesteps, self.max_batch = int(n_timesteps), int(max_batch)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("dynamics_aware_diffusion_b200 runs on CUDA (sm_100a) only; got device %s" % device)
        self.precision = precision
        cfg = N.DadConfig(abi_version=N.DAD_ABI_VERSION, device=self.device.index or 0,
                          precision=N.PRECISION_BF16 if precision == "bf16" else N.PRECISION_FP32,
                          transition_dim=transition_dim, dim=dim, n_levels=len(dim_mults), kernel_size=kernel_size,
                          time_dim=time_dim or 0, horizon=horizon, n_timesteps=n_timesteps,
                          predict_epsilon=int(bool(predict_epsilon)), clip_denoised=int(bool(clip_denoised)),
                          max_batch=max_batch)
        for i, m in enumerate(dim_mults):
            cfg.dim_mults[i] = int(m)
        rc = self.lib.dad_create(ctypes.byref(cfg), ctypes.byref(self.handle))
        if rc != 0:
            msg = self.lib.dad_last_error(None)
            self.handle = ctypes.c_void_p()
            raise N.DadError(rc, msg.decode() if msg else "dad_create failed")
        self._keep = []          # tensors whose device memory the handle may still read asynchronously
        self.has_projector = False
        self.projector_tag = None        # who pushed the projector that is loaded now (DynamicsAwarePolicy._push_projector)
        self._companion = None           # (fp32 Engine, min_step) attached with set_fp32_steps
        self.fp32_math = "fp32"

    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.dad_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        N.check(self.handle, rc)

    # ---- set-up ---------------------------------------------------------------------------------
    def load_unet_state(self, named):
        """named: iterable of (key relative to TemporalUnet, fp32 tensor on any device)."""
        items = [(k, v.detach().to(torch.float32).contiguous()) for k, v in named]
        arr = (N.DadTensor * len(items))()
        keep = []
        for i, (k, v) in enumerate(items):
            kb = k.encode()
            keep.append(kb)
            arr[i].name = kb
            arr[i].data = v.data_ptr()
            arr[i].numel = v.numel()
        torch.cuda.synchronize(self.device)      # sampling still in flight reads the tables this call rewrites
        self._ck(self.lib.dad_load_weights(self.handle, arr, len(items)))

    def set_schedule(self, sqrt_recip, sqrt_recipm1, coef1, coef2, log_var):
        ts = [t.detach().to("cpu", torch.float32).contiguous() for t in (sqrt_recip, sqrt_recipm1, coef1, coef2, log_var)]
        torch.cuda.synchronize(self.device)
        self._ck(self.lib.dad_set_schedule(self.handle, *[_ptr(t) for t in ts], ts[0].numel()))

    def set_projector(self, Nmat, q, alpha):
        torch.cuda.synchronize(self.device)
        self.projector_tag = None
        if Nmat is None:
            self._ck(self.lib.dad_set_projector(self.handle, None, None, None, 0, 0))
            self.has_projector = False
            return
        Nm = torch.as_tensor(Nmat, dtype=torch.float32).contiguous().cpu()
        qq = torch.as_tensor(q, dtype=torch.float32).contiguous().cpu()
        al = torch.as_tensor(alpha, dtype=torch.float32).contiguous().cpu()
        self._ck(self.lib.dad_set_projector(self.handle, _ptr(Nm), _ptr(qq), _ptr(al), Nm.shape[0], al.numel()))
        self.has_projector = True

    def set_conditions(self, conditions, batch_size):
        """conditions: dict h -> tensor broadcastable to (B, T) (GuidedPolicy.apply_conditions)."""
        if not conditions:
            self._ck(self.lib.dad_set_conditions(self.handle, None, None, 0, 0, 0))
            return 0
        T = self.transition_dim
        hs, vals, per_batch = [], [], False
        for h, v in conditions.items():
            v = torch.as_tensor(v, dtype=torch.float32)
            v = v.reshape(-1, T) if v.dim() <= 2 else v
            if v.shape[0] not in (1, batch_size):
                raise ValueError("condition at h=%s has batch %d, expected 1 or %d" % (h, v.shape[0], batch_size))
            per_batch |= v.shape[0] != 1
            hs.append(int(h))
            vals.append(v)
        if per_batch:
            vals = [v.expand(batch_size, T) for v in vals]
        buf = torch.stack([v.to(self.device) for v in vals]).contiguous()      # (n_cond, 1|B, T)
        idx = (ctypes.c_int32 * len(hs))(*hs)
        torch.cuda.synchronize(self.device)
        self._ck(self.lib.dad_set_conditions(self.handle, idx, _ptr(buf), len(hs), int(per_batch), batch_size))
        return len(hs)

    # ---- compute --------------------------------------------------------------------------------
    def unet_forward(self, x, t=None, step=0):
        x = _f32c(x, "x")
        B = x.shape[0]
        eps = torch.empty_like(x)
        if t is not None:
            t = t.to(device=x.device, dtype=torch.int64).contiguous()
        with torch.cuda.device(self.device):
            self._ck(self.lib.dad_unet_forward(self.handle, _ptr(x), _ptr(t), int(step), _ptr(eps), B, _stream()))
        self._keep = [x, t]
        return eps

    def step(self, x, model_out, step, noise=None, grad=None, guide_w=0.0, flags=0, seed=0, sample_offset=0):
        """In place on x."""
        x = _f32c(x, "x")
        model_out = _f32c(model_out, "model_out")
        if noise is not None:
            noise = _f32c(noise, "noise")
        if grad is not None:
            grad = _f32c(grad, "grad")
        with torch.cuda.device(self.device):
            self._ck(self.lib.dad_step(self.handle, _ptr(x), _ptr(model_out), _ptr(noise), _ptr(grad), float(guide_w),
                                       int(step), int(flags), int(seed), int(sample_offset), x.shape[0], _stream()))
        self._keep = [x, model_out, noise, grad]
        return x

    # ---- guided sampling inside a caller-captured CUDA graph (dad_loop_*) -----------------------------
    def loop_begin(self, x, n_steps, noise=None, noise_single=False, grad=None, guide_w=0.0, flags=0, seed=0,
                   sample_offset=0, trace=None):
        """Put the loop state on the device: x in/out, noise None (Philox) | (n_steps,B,H,T) | one reused (B,H,T)
        slot, grad = the buffer the caller's guidance writes every step."""
        x = _f32c(x, "x")
        with torch.cuda.device(self.device):
            self._ck(self.lib.dad_loop_begin(self.handle, _ptr(x), _ptr(noise), 1 if noise_single else 0, _ptr(grad),
                                             float(guide_w), int(seed), int(sample_offset), x.shape[0], int(n_steps),
                                             int(flags), _ptr(trace), _stream()))
        self._keep = [x, noise, grad, trace]

    def loop_unet(self, B):
        """Enqueue `step -= 1; eps = unet(x)` on the current stream (capturable)."""
        self._ck(self.lib.dad_loop_unet(self.handle, int(B), _stream()))

    def loop_step(self, B, flags=0):
        """Enqueue the fused remainder of the step (reads eps, grad, noise; writes x) on the current stream (capturable)."""
        self._ck(self.lib.dad_loop_step(self.handle, int(B), int(flags), _stream()))

    def graph_epoch(self):
        return int(self.lib.dad_graph_epoch(self.handle))

    def loop_replayed(self, n):
        self._ck(self.lib.dad_loop_replayed(self.handle, int(n)))

    def project(self, x, step):
        x = _f32c(x, "x")
        with torch.cuda.device(self.device):
            self._ck(self.lib.dad_project(self.handle, _ptr(x), int(step), x.shape[0], _stream()))
        return x

    def sample(self, x, n_steps, noise_seq=None, flags=0, seed=0, sample_offset=0, trace=None):
        """In place on x: x_S -> x_0 through n_steps graph-replayed steps."""
        x = _f32c(x, "x")
        if noise_seq is not None:
            noise_seq = _f32c(noise_seq, "noise_seq")
            if noise_seq.shape[0] < n_steps or noise_seq.shape[1:] != x.shape:
                raise ValueError("noise_seq must be (n_steps, B, H, T)")
        if trace is not None:
            trace = _f32c(trace, "trace")
        with torch.cuda.device(self.device):
            self._ck(self.lib.dad_sample(self.handle, _ptr(x), _ptr(noise_seq), int(seed), int(sample_offset),
                                         x.shape[0], int(n_steps), int(flags), _ptr(trace), _stream()))
        self._keep = [x, noise_seq, trace]
        return x

    def sample_host(self, x_host, n_steps, noise_seq_host=None, flags=0, seed=0, sample_offset=0):
        """Host (CPU, ideally pinned) buffers in and out; synchronous.  x_host is overwritten with x_0."""
        if x_host.is_cuda or x_host.dtype != torch.float32 or not x_host.is_contiguous():
            raise ValueError("x_host must be a contiguous float32 CPU tensor")
        self._ck(self.lib.dad_sample_host(self.handle, _ptr(x_host), _ptr(noise_seq_host), int(seed),
                                          int(sample_offset), x_host.shape[0], int(n_steps), int(flags)))
        return x_host

    # ---- measurement hooks ----------------------------------------------------------------------
    def sample_profile(self, x, n_steps, flags=0, seed=0, sample_offset=0):
        """sample() with Philox noise, returning per-diffusion-step device times in ms (list of n_steps)."""
        x = _f32c(x, "x")
        buf = (ctypes.c_float * int(n_steps))()
        with torch.cuda.device(self.device):
            self._ck(self.lib.dad_sample_profile(self.handle, _ptr(x), int(seed), int(sample_offset), x.shape[0],
                                                 int(n_steps), int(flags), buf, _stream()))
        return list(buf)

    def layers(self):
        out = []
        for i in range(self.lib.dad_layer_count(self.handle)):
            d = N.DadLayerDesc()
            self._ck(self.lib.dad_layer_info(self.handle, i, ctypes.byref(d)))
            out.append({"index": i, "name": d.name.decode(), "L_out": d.L_out, "C_in": d.C_in, "C_out": d.C_out,
                        "taps": d.taps, "tile_n": d.tile_n, "group_width": d.group_width,
                        "flops_per_sample": d.flops_per_sample, "kernel": d.kernel.decode()})
        return out

    def time_layer(self, index, B, iters=20):
        ms = ctypes.c_float()
        with torch.cuda.device(self.device):
            self._ck(self.lib.dad_time_layer(self.handle, int(index), int(B), int(iters), ctypes.byref(ms), _stream()))
        return ms.value

    def units(self):
        """Launch units of one U-Net pass at the current fusion level (a conv chain or a single layer)."""
        out = []
        for i in range(self.lib.dad_unit_count(self.handle)):
            d = N.DadUnitDesc()
            self._ck(self.lib.dad_unit_info(self.handle, i, ctypes.byref(d)))
            out.append({"index": i, "first_layer": d.first_layer, "n_layers": d.n_layers, "is_chain": bool(d.is_chain),
                        "L_out": d.L_out, "C_out": d.C_out, "flops_per_sample": d.flops_per_sample,
                        "kernel": d.kernel.decode()})
        return out

    def time_unit(self, index, B, iters=20):
        ms = ctypes.c_float()
        with torch.cuda.device(self.device):
            self._ck(self.lib.dad_time_unit(self.handle, int(index), int(B), int(iters), ctypes.byref(ms), _stream()))
        return ms.value

    def debug_counters(self, reset=False):
        """(error code, waits that spun, ns spent spinning, all dependency waits) of the conv chains; the last three
        are counted by -DDAD_TUNING builds only."""
        out = (ctypes.c_uint32 * 4)()
        self._ck(self.lib.dad_debug_counters(self.handle, ctypes.byref(out), 1 if reset else 0))
        return tuple(int(v) for v in out)

    def set_fusion(self, level):
        """0 = per-layer kernels of round 1, 1 = chain kernel per conv, 2 = per ResidualTemporalBlock, 3 = per run of
        blocks (default)."""
        self._ck(self.lib.dad_set_fusion(self.handle, int(level)))

    def time_step_kernel(self, B, step, flags=0, iters=20):
        ms = ctypes.c_float()
        with torch.cuda.device(self.device):
            self._ck(self.lib.dad_time_step_kernel(self.handle, int(B), int(step), int(flags), int(iters),
                                                   ctypes.byref(ms), _stream()))
        return ms.value

    def info(self):
        out = N.DadInfo()
        self._ck(self.lib.dad_get_info(self.handle, ctypes.byref(out)))
        return {f: getattr(out, f) for f, _ in N.DadInfo._fields_}

    def launch_count(self):
        return int(self.lib.dad_launch_count(self.handle))

    def set_fp32_steps(self, companion, min_step=0):
        """Reverse steps with index >= min_step take eps from `companion` (an fp32 Engine of the same architecture and
        weights) inside sample() / sample_host(); None detaches.  This engine keeps the companion alive."""
        if companion is None:
            self._ck(self.lib.dad_set_fp32_steps(self.handle, None, 0))
            self._companion = None
            return
        self._ck(self.lib.dad_set_fp32_steps(self.handle, companion.handle, int(min_step)))
        self._companion = (companion, int(min_step))

    FP32_MATH = {"fp32": 0, "tf32": 1, "tf32x3": 2}

    def set_fp32_math(self, mode):
        """fp32-precision engines: 'fp32' IEEE SIMT (default, the 1e-5 parity mode), 'tf32' TF32 operands on the tensor
        cores (fp32 accumulate, fp32 activations), 'tf32x3' 3xTF32 error compensation."""
        self._ck(self.lib.dad_set_fp32_math(self.handle, self.FP32_MATH[mode]))
        self.fp32_math = mode

    def set_latency_batch(self, max_b):
        """Batches <= max_b run the latency kernels (get_action's single plan); 0 = throughput kernels only."""
        self._ck(self.lib.dad_set_latency_batch(self.handle, int(max_b)))


def create_engine_auto(precision, **kw):
    """precision 'auto': bf16 tensor-core path when the architecture fits it, else the fp32 CUDA path."""
    if precision in ("bf16", "fp32"):
        return Engine(precision=precision, **kw)
    try:
        return Engine(precision="bf16", **kw)
    except N.DadError as e:
        if e.code != N.ERR_INVALID:
            raise
        warnings.warn("bf16 tensor-core path unavailable for this architecture (%s); using the fp32 CUDA path" % e)
        return Engine(precision="fp32", **kw)
