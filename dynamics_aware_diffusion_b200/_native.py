"""ctypes binding of include/dad_b200.h.  There is no fallback: a missing library or a failed call raises."""
import ctypes
import os

from .build import LIB_PATH as _DEFAULT_LIB

# The in-tree build is THE library.  A tuning variant (build.py --out=...) is picked up only when DAD_TUNING=1 is set
# alongside DAD_LIB_PATH, so a stray variable cannot swap the kernels under a production process.
LIB_PATH = (os.environ.get("DAD_LIB_PATH", _DEFAULT_LIB) if os.environ.get("DAD_TUNING") == "1" else _DEFAULT_LIB)

DAD_ABI_VERSION = 1
DAD_MAX_LEVELS = 8
PRECISION_FP32, PRECISION_BF16 = 0, 1
FLAG_CONDITIONS, FLAG_PROJECT, FLAG_PROJECT_AFTER_INPAINT, FLAG_PHILOX_INIT = 1, 2, 4, 8
ERR_INVALID = -1

EXPORTS = (
    "dad_abi_version", "dad_create", "dad_destroy", "dad_last_error", "dad_load_weights",
    "dad_set_schedule", "dad_set_projector", "dad_set_conditions", "dad_unet_forward", "dad_step",
    "dad_project", "dad_sample", "dad_sample_host", "dad_get_info", "dad_launch_count", "dad_set_latency_batch",
    "dad_loop_begin", "dad_loop_unet", "dad_loop_step", "dad_graph_epoch", "dad_loop_replayed",
    "dad_build_projection_matrix", "dad_fit_linear_dynamics", "dad_dynamics_residual",
    "dad_sample_profile", "dad_layer_count", "dad_layer_info", "dad_time_layer", "dad_time_step_kernel",
    "dad_set_fusion", "dad_unit_count", "dad_unit_info", "dad_time_unit", "dad_debug_counters", "dad_set_fp32_steps", "dad_set_fp32_math",
)


class DadConfig(ctypes.Structure):
    _fields_ = [
        ("abi_version", ctypes.c_int32), ("device", ctypes.c_int32), ("precision", ctypes.c_int32),
        ("transition_dim", ctypes.c_int32), ("dim", ctypes.c_int32), ("n_levels", ctypes.c_int32),
        ("dim_mults", ctypes.c_int32 * DAD_MAX_LEVELS), ("kernel_size", ctypes.c_int32),
        ("time_dim", ctypes.c_int32), ("horizon", ctypes.c_int32), ("n_timesteps", ctypes.c_int32),
        ("predict_epsilon", ctypes.c_int32), ("clip_denoised", ctypes.c_int32), ("max_batch", ctypes.c_int32),
    ]


class DadTensor(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("data", ctypes.c_void_p), ("numel", ctypes.c_int64)]


class DadInfo(ctypes.Structure):
    _fields_ = [("conv_flops_per_sample", ctypes.c_int64), ("launches_per_step", ctypes.c_int64),
                ("workspace_bytes", ctypes.c_int64), ("n_conv_layers", ctypes.c_int32), ("sm_count", ctypes.c_int32)]


class DadLayerDesc(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 96), ("L_out", ctypes.c_int32), ("C_in", ctypes.c_int32),
                ("C_out", ctypes.c_int32), ("taps", ctypes.c_int32), ("tile_n", ctypes.c_int32),
                ("group_width", ctypes.c_int32), ("flops_per_sample", ctypes.c_int64), ("kernel", ctypes.c_char * 64)]


class DadUnitDesc(ctypes.Structure):
    _fields_ = [("first_layer", ctypes.c_int32), ("n_layers", ctypes.c_int32), ("is_chain", ctypes.c_int32),
                ("L_out", ctypes.c_int32), ("C_out", ctypes.c_int32), ("flops_per_sample", ctypes.c_int64),
                ("kernel", ctypes.c_char * 64)]


class DadError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("dad_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    """The loaded C-ABI library.  Raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "CUDA library %s is missing: build it with `python -m dynamics_aware_diffusion_b200.build` "
            "(or __graft_entry__.build()).  This package has no CPU or PyTorch fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, u32, u64, f32 = (ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32,
                                   ctypes.c_uint64, ctypes.c_float)
    L.dad_abi_version.restype = ctypes.c_int
    L.dad_abi_version.argtypes = []
    L.dad_create.argtypes = [ctypes.POINTER(DadConfig), ctypes.POINTER(vp)]
    L.dad_destroy.argtypes = [vp]
    L.dad_last_error.argtypes = [vp]
    L.dad_last_error.restype = ctypes.c_char_p
    L.dad_load_weights.argtypes = [vp, ctypes.POINTER(DadTensor), i32]
    L.dad_set_schedule.argtypes = [vp, vp, vp, vp, vp, vp, i32]
    L.dad_set_projector.argtypes = [vp, vp, vp, vp, i32, i32]
    L.dad_set_conditions.argtypes = [vp, vp, vp, i32, i32, i32]
    L.dad_unet_forward.argtypes = [vp, vp, vp, i32, vp, i32, vp]
    L.dad_step.argtypes = [vp, vp, vp, vp, vp, f32, i32, u32, u64, u64, i32, vp]
    L.dad_project.argtypes = [vp, vp, i32, i32, vp]
    L.dad_sample.argtypes = [vp, vp, vp, u64, u64, i32, i32, u32, vp, vp]
    L.dad_sample_host.argtypes = [vp, vp, vp, u64, u64, i32, i32, u32]
    L.dad_get_info.argtypes = [vp, ctypes.POINTER(DadInfo)]
    L.dad_launch_count.argtypes = [vp]
    L.dad_launch_count.restype = ctypes.c_int64
    L.dad_set_latency_batch.argtypes = [vp, ctypes.c_int32]
    L.dad_set_fp32_steps.argtypes = [vp, vp, ctypes.c_int32]
    L.dad_set_fp32_math.argtypes = [vp, ctypes.c_int32]
    L.dad_loop_begin.argtypes = [vp, vp, vp, ctypes.c_int32, vp, ctypes.c_float, ctypes.c_uint64, ctypes.c_uint64,
                                 ctypes.c_int32, ctypes.c_int32, ctypes.c_uint32, vp, vp]
    L.dad_loop_unet.argtypes = [vp, ctypes.c_int32, vp]
    L.dad_loop_step.argtypes = [vp, ctypes.c_int32, ctypes.c_uint32, vp]
    L.dad_graph_epoch.argtypes = [vp]
    L.dad_graph_epoch.restype = ctypes.c_int64
    L.dad_loop_replayed.argtypes = [vp, ctypes.c_int32]
    L.dad_build_projection_matrix.argtypes = [ctypes.c_int32, vp, ctypes.c_int32, ctypes.c_int32, vp]
    L.dad_fit_linear_dynamics.argtypes = [ctypes.c_int32, vp, vp, vp, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, vp, vp]
    L.dad_dynamics_residual.argtypes = [ctypes.c_int32, vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                        vp, vp, vp, vp, vp, ctypes.POINTER(ctypes.c_double), vp]
    L.dad_sample_profile.argtypes = [vp, vp, u64, u64, i32, i32, u32, vp, vp]
    L.dad_layer_count.argtypes = [vp]
    L.dad_layer_info.argtypes = [vp, i32, ctypes.POINTER(DadLayerDesc)]
    L.dad_time_layer.argtypes = [vp, i32, i32, i32, ctypes.POINTER(ctypes.c_float), vp]
    L.dad_time_step_kernel.argtypes = [vp, i32, i32, u32, i32, ctypes.POINTER(ctypes.c_float), vp]
    L.dad_debug_counters.argtypes = [vp, ctypes.POINTER(ctypes.c_uint32 * 4), i32]
    L.dad_set_fusion.argtypes = [vp, i32]
    L.dad_unit_count.argtypes = [vp]
    L.dad_unit_info.argtypes = [vp, i32, ctypes.POINTER(DadUnitDesc)]
    L.dad_time_unit.argtypes = [vp, i32, i32, i32, ctypes.POINTER(ctypes.c_float), vp]
    if L.dad_abi_version() != DAD_ABI_VERSION:
        raise RuntimeError("libdad_b200.so ABI %d != binding ABI %d; rebuild" % (L.dad_abi_version(), DAD_ABI_VERSION))
    _lib = L
    return L


def check(handle, code):
    if code != 0:
        msg = lib().dad_last_error(handle)
        raise DadError(code, msg.decode() if msg else "unknown")
