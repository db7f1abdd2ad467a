"""ctypes binding of include/evpp_b200.h.  The product path: no CUDA library -> hard error."""
import ctypes as C
import os

_LIB = None
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libevpp_b200.so")

EVPP_PREC_FP32, EVPP_PREC_FP16, EVPP_PREC_BF16, EVPP_PREC_FP16X2, EVPP_PREC_FP16X3 = 0, 1, 2, 3, 4
PRECISIONS = {"fp32": 0, "fp16": 1, "bf16": 2, "fp16x2": 3, "fp16x3": 4}
EVPP_ERR_NONFINITE = -5


class EvppConfig(C.Structure):
    _fields_ = [("n_samples", C.c_int32), ("n_importance", C.c_int32), ("near_plane", C.c_float),
                ("far_plane", C.c_float), ("white_bkgd", C.c_int32), ("lindisp", C.c_int32),
                ("precision", C.c_int32), ("max_chunk_rays", C.c_int32)]


class EvppRenderOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("rgb", "disp", "acc", "depth", "unc", "raw", "rgb0", "disp0",
                                          "acc0", "depth0", "unc0", "z_std")]


class EvppNgpConfig(C.Structure):
    _fields_ = [("aabb_min", C.c_float * 3), ("aabb_max", C.c_float * 3), ("grid_res", C.c_int32),
                ("near_plane", C.c_float), ("far_plane", C.c_float), ("dt", C.c_float), ("max_steps", C.c_int32),
                ("max_samples", C.c_int32), ("n_levels", C.c_int32), ("white_bkgd", C.c_int32),
                ("march_mode", C.c_int32), ("cascades", C.c_int32), ("dt_gamma", C.c_float), ("min_near", C.c_float)]


class EvppNgpOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("rgb", "depth", "acc", "beta2", "counts", "t", "raw", "enc")] + \
               [("sample_capacity", C.c_int64), ("total_so_far", C.c_int64), ("delta", C.c_void_p), ("cell", C.c_void_p)]


class EvppError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"evpp_b200 error {code}: {msg}")
        self.code = code


# every symbol include/evpp_b200.h declares: name -> (restype, argtypes)
_P, _I, _F, _L = C.c_void_p, C.c_int, C.c_float, C.c_int64
SIGNATURES = {
    "evpp_last_error": (C.c_char_p, []),
    "evpp_abi_version": (_I, []),
    "evpp_device_count": (_I, []),
    "evpp_create": (_I, [_I, C.POINTER(EvppConfig), C.POINTER(_P)]),
    "evpp_destroy": (_I, [_P]),
    "evpp_load_nerf_weights": (_I, [_P, _I, _P, _P, _I]),
    "evpp_score_views": (_I, [_P, _P, _I, _I, _I, _F, _F, _F, _F, _P, _I, _P, _P]),
    "evpp_score_views_ex": (_I, [_P, _P, _I, _I, _I, _F, _F, _F, _F, _P, _I, _I, _P, _P]),
    "evpp_get_nonfinite": (_I, [_P, C.POINTER(C.c_int32), _P]),
    "evpp_score_views_host": (_I, [_P, _P, _I, _I, _I, _F, _F, _F, _F, _P, _I, _P, _P]),
    "evpp_rescore_views_fp32": (_I, [_P, _P, _P, _I, _I, _I, _F, _F, _F, _F, _P, _I, _P, _P]),
    "evpp_rescore_views": (_I, [_P, _P, _P, _I, _I, _I, _F, _F, _F, _F, _P, _I, _I, _P, _P]),
    "evpp_set_score_mode": (_I, [_P, _I]),
    "evpp_set_render_mode": (_I, [_P, _I]),
    "evpp_render_rays": (_I, [_P, _P, _P, _L, _F, _F, _I, C.POINTER(EvppRenderOut), _P]),
    "evpp_render_rays_ex": (_I, [_P, _P, _P, _P, _L, _F, _F, _I, C.POINTER(EvppRenderOut), _P]),
    "evpp_get_rays": (_I, [_P, _P, _I, _I, _I, _F, _F, _F, _F, _P, _I, _P, _P, _P, _P]),
    "evpp_mlp_forward": (_I, [_P, _I, _I, _P, _P, _L, _P, _P]),
    "evpp_mlp_layer_dump": (_I, [_P, _I, _I, _P, _P, _L, _I, _P, _P, _P]),
    "evpp_mlp_trace": (_I, [_P, _I, _I, _P, _P, _L, _P, _P, _P]),
    "evpp_set_stage_capture": (_I, [_P, _I]),
    "evpp_stage_dump": (_I, [_P, _I, _P, C.POINTER(_L), _P]),
    "evpp_composite": (_I, [_P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "evpp_resample": (_I, [_P, _P, _P, _P, _I, _P, _P, _P, _P, _P]),
    "evpp_select_nbv": (_I, [_P, _P, _P, _I, _I, _P, _P, _P, _P]),
    "evpp_sample_pixels": (_I, [_P, C.c_uint64, _I, _I, _I, _P, _P]),
    "evpp_surrogate_param_count": (_I, []),
    "evpp_surrogate_fit": (_I, [_P, _P, _P, _P, _I, _I, _F, _P, _P]),
    "evpp_surrogate_query": (_I, [_P, _P, _P, _L, _P, _P]),
    "evpp_tsdf_set_volume": (_I, [_P, _P, _P, _P, _P, _F, _P, _P, _F, _F, _I, _F, _F]),
    "evpp_tsdf_render_rays": (_I, [_P, _P, _P, _L, _P, _P, _P, _P, _P, _P]),
    "evpp_tsdf_score_views": (_I, [_P, _P, _I, _I, _I, _F, _F, _F, _F, _P, _I, _P, _P, _P]),
    "evpp_tsdf_score_views_host": (_I, [_P, _P, _I, _I, _I, _F, _F, _F, _F, _P, _I, _P, _P, _P]),
    "evpp_ngp_configure": (_I, [_P, C.POINTER(EvppNgpConfig), _P, _P, _P, _P]),
    "evpp_ngp_set_occupancy": (_I, [_P, _P, _L]),
    "evpp_ngp_load_params": (_I, [_P, _P, _L, _P, _L]),
    "evpp_tensorf_load_params": (_I, [_P, _P, _P, _I, _P, _L]),
    "evpp_ngp_encode": (_I, [_P, _P, _L, _P, _P, _P]),
    "evpp_ngp_render_rays": (_I, [_P, _P, _P, _L, C.POINTER(EvppNgpOut), _P]),
    "evpp_ngp_score_views": (_I, [_P, _P, _I, _I, _I, _F, _F, _F, _F, _P, _I, _P, _P]),
    "evpp_ngp_score_views_host": (_I, [_P, _P, _I, _I, _I, _F, _F, _F, _F, _P, _I, _P, _P]),
    "evpp_ngp_get_sample_count": (_I, [_P, C.POINTER(_L)]),
    "evpp_check_guards": (_I, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "evpp_poison_workspace": (_I, [_P]),
    "evpp_get_counters": (_I, [_P, C.POINTER(_L), C.POINTER(_L)]),
    "evpp_set_profiling": (_I, [_P, _I]),
    "evpp_get_mlp_time_ms": (_I, [_P, C.POINTER(C.c_double), C.POINTER(_L), C.POINTER(_L)]),
    "evpp_get_mlp_profile": (_I, [_P, C.POINTER(C.c_double), C.POINTER(_L), C.POINTER(_L), C.POINTER(C.c_double)]),
}


def load():
    """Load the CUDA library; raises if it has not been built (no CPU fallback exists)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise EvppError(-4, f"{LIB_PATH} not built; run `python -m evpp_b200.build` "
                            "(evpp_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().evpp_last_error()
        raise EvppError(rc, msg.decode() if msg else "unknown")
