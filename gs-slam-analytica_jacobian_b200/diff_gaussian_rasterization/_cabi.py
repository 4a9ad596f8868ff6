"""ctypes binding of libgsr_b200.so (C-ABI declared in include/gsr_b200.h).

This replaces the reference's pybind module `diff_gaussian_rasterization._C` (ext.cpp:15-19).
There is NO fallback: if the CUDA library is missing or fails to load, importing the operator
raises immediately.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgsr_b200.so")

GSR_OK = 0
GSR_ERR_ARG, GSR_ERR_CUDA, GSR_ERR_WORKSPACE, GSR_ERR_OVERFLOW, GSR_ERR_TIMEOUT = -1, -2, -3, -4, -5

# every symbol include/gsr_b200.h declares (checked by tests/test_cabi_symbols.py)
EXPORTS = (
    "gsr_geometry_bytes", "gsr_image_bytes", "gsr_binning_bytes", "gsr_forward_plan", "gsr_forward_num_rendered",
    "gsr_forward_render", "gsr_forward_overflowed", "gsr_rasterize_gaussians", "gsr_rasterize_gaussians_backward",
    "gsr_mark_visible", "gsr_debug_pointers", "gsr_error_string", "gsr_version", "gsr_kernel_launch_count",
    "gsr_stage_timing", "gsr_stage_times_ms", "gsr_debug_probe", "gsr_slam_loss_scratch_bytes", "gsr_slam_loss",
    "gsr_tracking_step", "gsr_forward_nosync", "gsr_forward_nosync_fuses_scatter", "gsr_sort_on_demand", "gsr_fused_loss_scratch_bytes",
    "gsr_step_status", "gsr_window_allreduce", "gsr_spatial_order",
)


class GsrFusedLoss(C.Structure):
    _fields_ = [
        ("gt_color", C.c_void_p), ("gt_depth", C.c_void_p), ("grad_mask", C.c_void_p), ("exposure", C.c_void_p),
        ("rgb_boundary_threshold", C.c_float), ("alpha", C.c_float), ("use_depth", C.c_int), ("opacity_weighted", C.c_int),
        ("dL_dcolor", C.c_void_p), ("dL_ddepth", C.c_void_p), ("sums", C.c_void_p), ("scratch", C.c_void_p),
    ]


class GsrScene(C.Structure):
    _fields_ = [
        ("P", C.c_int), ("D", C.c_int), ("M", C.c_int), ("W", C.c_int), ("H", C.c_int),
        ("background", C.c_void_p), ("means3D", C.c_void_p), ("shs", C.c_void_p), ("colors_precomp", C.c_void_p),
        ("opacities", C.c_void_p), ("scales", C.c_void_p), ("rotations", C.c_void_p), ("cov3D_precomp", C.c_void_p),
        ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("projmatrix_raw", C.c_void_p), ("campos", C.c_void_p),
        ("scale_modifier", C.c_float), ("tan_fovx", C.c_float), ("tan_fovy", C.c_float),
        ("prefiltered", C.c_int), ("debug", C.c_int), ("accumulate_grads", C.c_int),
        ("densify_grad_accum", C.c_void_p), ("densify_denom", C.c_void_p), ("max_radii2D", C.c_void_p),
        ("overlap_forward", C.c_int), ("upstream_ready", C.c_void_p), ("fused_loss", C.POINTER(GsrFusedLoss)),
        ("sort_on_demand", C.c_int), ("exact_exp", C.c_int), ("tile_row_begin", C.c_int), ("tile_row_end", C.c_int), ("spatial_order", C.c_void_p), ("depth_cut", C.c_void_p),
    ]


ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)

_lib = None


def _build_if_possible():
    import importlib.util

    spec = importlib.util.spec_from_file_location("gsr_b200_build", os.path.join(os.path.dirname(_HERE), "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()


def load():
    """Load (building first when the sources are newer and nvcc is available) the CUDA library."""
    global _lib
    if _lib is not None:
        return _lib
    try:
        _build_if_possible()
    except Exception as ex:  # stale/missing library and no compiler: fail loudly below if it is missing
        if not os.path.exists(LIB_PATH):
            raise ImportError("gsr_b200: CUDA library %s is missing and could not be built: %s" % (LIB_PATH, ex))
    if not os.path.exists(LIB_PATH):
        raise ImportError("gsr_b200: CUDA library %s is missing (run gs-slam-analytica_jacobian_b200/build.py)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, sz, ll, ip = C.c_void_p, C.c_size_t, C.c_longlong, C.c_int
    sp = C.POINTER(GsrScene)
    lib.gsr_geometry_bytes.restype = sz
    lib.gsr_geometry_bytes.argtypes = [ip, ip, ip]
    lib.gsr_image_bytes.restype = sz
    lib.gsr_image_bytes.argtypes = [ip, ip]
    lib.gsr_binning_bytes.restype = sz
    lib.gsr_binning_bytes.argtypes = [ip, ip, ip, ll]
    lib.gsr_forward_plan.argtypes = [sp, vp, sz, vp, vp, vp]
    lib.gsr_forward_num_rendered.argtypes = [vp, vp, C.POINTER(ll), C.POINTER(ll)]
    lib.gsr_forward_render.argtypes = [sp, vp, vp, sz, ll, ll, ll, vp, sz, vp, vp, vp, vp, vp]
    lib.gsr_forward_nosync.argtypes = [sp, vp, sz, vp, sz, ll, ll, vp, sz, vp, vp, vp, vp, vp, vp]
    lib.gsr_forward_nosync_fuses_scatter.argtypes = [ip, ip, ip]
    lib.gsr_forward_overflowed.argtypes = [vp, vp, C.POINTER(ip), C.POINTER(ll)]
    lib.gsr_step_status.argtypes = [vp, vp, C.POINTER(C.c_uint)]
    lib.gsr_rasterize_gaussians.argtypes = [sp, vp, sz, vp, sz, ALLOC_FN, vp, C.POINTER(vp), C.POINTER(ll),
                                            vp, vp, vp, vp, vp, vp]
    lib.gsr_rasterize_gaussians_backward.argtypes = [sp, vp, vp, vp, ll, vp, vp, vp] + [vp] * 9 + [vp]
    lib.gsr_mark_visible.argtypes = [ip, vp, vp, vp, vp, vp]
    lib.gsr_debug_pointers.argtypes = [ip, ip, ip, vp, vp, ll, vp, C.POINTER(C.c_ulonglong)]
    lib.gsr_error_string.restype = C.c_char_p
    lib.gsr_kernel_launch_count.restype = C.c_ulonglong
    lib.gsr_stage_timing.argtypes = [ip]
    lib.gsr_sort_on_demand.argtypes = [ip]
    lib.gsr_fused_loss_scratch_bytes.argtypes = [ip, ip]
    lib.gsr_fused_loss_scratch_bytes.restype = sz
    lib.gsr_stage_times_ms.argtypes = [C.POINTER(C.c_float)]
    lib.gsr_debug_probe.argtypes = [C.POINTER(C.c_ulonglong), sz]
    lib.gsr_slam_loss_scratch_bytes.restype = sz
    lib.gsr_slam_loss_scratch_bytes.argtypes = [ip, ip]
    fl = C.c_float
    lib.gsr_slam_loss.argtypes = [ip, ip, vp, vp, vp, vp, vp, vp, vp, fl, fl, ip, ip, vp, vp, vp, vp, vp]
    lib.gsr_tracking_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, fl, fl, fl, fl, vp]
    lib.gsr_spatial_order.argtypes = [sp, vp, sz, vp, vp]
    lib.gsr_window_allreduce.argtypes = [vp, vp, ip, ip, sz, ip, sz, vp, vp]
    for name in EXPORTS:
        getattr(lib, name)
    _lib = lib
    return lib


def check(rc, what):
    if rc != GSR_OK:
        msg = load().gsr_error_string().decode("utf-8", "replace")
        if rc == GSR_ERR_ARG:
            raise Exception(msg)          # the reference raises plain Exception for bad argument combos
        raise RuntimeError("gsr_b200 %s failed (%d): %s" % (what, rc, msg))
