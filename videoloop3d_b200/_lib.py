"""ctypes binding of libvl3d.so (C ABI in include/vl3d.h).

There is deliberately NO fallback: if the CUDA library is missing or fails to load, every product
entry point raises.  (The CPU oracle lives under oracle/ and is test infrastructure only.)
"""
import ctypes as C
import os

from . import build as _build

MAX_PLANES = 32


class Quad(C.Structure):
    _fields_ = [("x0f", C.c_float), ("y0f", C.c_float), ("sx", C.c_float), ("sy", C.c_float),
                ("x0i", C.c_int32), ("y0i", C.c_int32), ("kind", C.c_int32), ("reserved", C.c_int32)]


class View(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("D", C.c_int32), ("qh", C.c_int32), ("qw", C.c_int32),
                ("dyn_h", C.c_int32), ("dyn_w", C.c_int32), ("sta_h", C.c_int32), ("sta_w", C.c_int32),
                ("cx", C.c_float), ("cy", C.c_float), ("hom", C.c_float * (MAX_PLANES * 9)), ("flags", C.c_int32)]


VIEW_RECT_PLANES = 1


class Own(C.Structure):
    """vl3d_own: per-plane texel -> screen maps for the owner mode of the fused backward + Adam."""
    _fields_ = [("hinv", C.c_float * (MAX_PLANES * 9)), ("reach", C.c_float * MAX_PLANES), ("rect", C.c_int32 * (MAX_PLANES * 4))]


class LossDesc(C.Structure):
    _fields_ = [("t", C.c_int32), ("F", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("p", C.c_int32), ("pt", C.c_int32), ("s", C.c_int32), ("st", C.c_int32),
                ("n1", C.c_int32), ("n2", C.c_int32), ("ho", C.c_int32), ("wo", C.c_int32),
                ("x_sf", C.c_int64), ("x_sc", C.c_int64), ("x_sr", C.c_int64),
                ("y_sf", C.c_int64), ("y_sc", C.c_int64), ("y_sr", C.c_int64),
                ("use_alpha", C.c_int32), ("alpha", C.c_float)]


class Box(C.Structure):
    """vl3d_box: one strided (frames, planes, rows, cols) copy of vl3d_copy_boxes."""
    _fields_ = [("src", C.c_void_p), ("src2", C.c_void_p), ("dst", C.c_void_p),
                ("n_frames", C.c_int32), ("n_planes", C.c_int32), ("n_rows", C.c_int32), ("n_cols", C.c_int32),
                ("src_sf", C.c_int64), ("src_sp", C.c_int64), ("src_sr", C.c_int64),
                ("dst_sf", C.c_int64), ("dst_sp", C.c_int64), ("dst_sr", C.c_int64)]


MAX_BOXES = 32
_P = C.c_void_p
_SIGNATURES = {
    "vl3d_version": (C.c_int, []),
    "vl3d_last_error_string": (C.c_char_p, []),
    "vl3d_composite_fwd": (C.c_int, [C.POINTER(View), _P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P]),
    "vl3d_composite_bwd": (C.c_int, [C.POINTER(View), _P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P]),
    "vl3d_composite_terms_fwd": (C.c_int, [C.POINTER(View), _P, _P, _P, _P, C.c_int32, _P, C.c_float, _P, _P, _P, _P]),
    "vl3d_composite_terms_bwd": (C.c_int, [C.POINTER(View), _P, _P, _P, _P, C.c_int32, _P, C.c_float, _P, _P, _P, _P, _P, _P]),
    "vl3d_scale_partials": (C.c_int, []),
    "vl3d_scale_invariant": (C.c_int, [_P, C.c_int32, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P]),
    "vl3d_frame_sum": (C.c_int, [_P, C.c_int32, C.c_int64, _P, _P]),
    "vl3d_scale_invariant_presum": (C.c_int, [_P, C.c_int32, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P]),
    "vl3d_scale_log_sum": (C.c_int, [_P, C.c_int32, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P]),
    "vl3d_scale_finish": (C.c_int, [_P, C.c_int64, _P, _P]),
    "vl3d_patchnn_search": (C.c_int, [C.POINTER(LossDesc), _P, _P, C.c_int32, C.c_int32, _P, _P]),
    "vl3d_scale_video": (C.c_int, [_P, _P, _P, C.c_int64, _P]),
    "vl3d_vote_partials": (C.c_int, [C.c_int32, C.c_int32, C.c_int32]),
    "vl3d_vote_loss": (C.c_int, [C.POINTER(LossDesc), _P, _P, _P, _P, C.c_int32, C.c_float, C.c_float, C.c_float,
                                 C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                 _P, _P, _P, _P, _P, _P]),
    "vl3d_video_loss_partials": (C.c_int, []),
    "vl3d_video_loss": (C.c_int, [C.c_int32, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "vl3d_patch_l1": (C.c_int, [C.POINTER(LossDesc), _P, _P, _P, _P, _P]),
    "vl3d_to8b": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "vl3d_u8_to_unit": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64, _P]),
    "vl3d_adam_step": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, _P]),
    "vl3d_fused_bwd_adam": (C.c_int, [C.POINTER(View), _P, _P, _P, C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int32,
                                      C.c_float, C.c_float, C.c_float, C.c_float, _P, C.c_int32, C.c_int32, _P, C.c_int32,
                                      _P, C.c_int32, _P]),
    "vl3d_copy_boxes": (C.c_int, [C.POINTER(Box), C.c_int32, _P]),
    "vl3d_alloc_compressible": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "vl3d_free_compressible": (C.c_int, [_P]),
    "vl3d_fused_own_scratch_bytes": (C.c_int64, [C.c_int32]),
    "vl3d_fused_own_table_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "vl3d_fused_bwd_adam_own": (C.c_int, [C.POINTER(View), _P, _P, _P, C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int32,
                                          C.c_float, C.c_float, C.c_float, C.c_float, _P, C.c_int32, C.c_int32, _P, C.c_int32,
                                          _P, C.c_int32, C.POINTER(Own), _P, C.c_int64, _P, C.c_int64, _P]),
}
EXPORTS = tuple(_SIGNATURES)

_lib = None
LAUNCHES = 0          # kernels launched through this binding (bench.py's gpu_launches)
_LAUNCHES_PER_CALL = {"vl3d_composite_fwd": 1, "vl3d_composite_bwd": 1, "vl3d_composite_terms_fwd": 1, "vl3d_composite_terms_bwd": 1, "vl3d_scale_invariant": 2, "vl3d_frame_sum": 1, "vl3d_scale_invariant_presum": 2, "vl3d_scale_log_sum": 2, "vl3d_scale_finish": 1,
                      "vl3d_patchnn_search": 1, "vl3d_scale_video": 1, "vl3d_patch_l1": 1, "vl3d_to8b": 1, "vl3d_u8_to_unit": 1, "vl3d_vote_loss": 2, "vl3d_video_loss": 2, "vl3d_adam_step": 1, "vl3d_fused_bwd_adam": 1, "vl3d_fused_bwd_adam_own": 2, "vl3d_copy_boxes": 1}


class Vl3dError(RuntimeError):
    pass


def lib_path():
    """In-tree library; VL3D_LIB=/path/to/other.so loads a different build of the same ABI (A/B measurements)."""
    return os.environ.get("VL3D_LIB") or _build.LIB_PATH


def load():
    """Load libvl3d.so (built in-tree by videoloop3d_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.isfile(path):
        raise Vl3dError(f"{path} not found: build it with `python -m videoloop3d_b200.build` "
                        f"(there is no CPU fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header / library mismatch
        fn.restype, fn.argtypes = res, args
    if lib.vl3d_version() != 100:
        raise Vl3dError(f"libvl3d version {lib.vl3d_version()} != 100")
    _lib = lib
    return lib


def call(name, *args):
    """Call an int-returning entry point and raise Vl3dError on a non-zero status."""
    global LAUNCHES
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.vl3d_last_error_string().decode(errors="replace")
        raise Vl3dError(f"{name} failed ({rc}): {msg}")
    LAUNCHES += _LAUNCHES_PER_CALL.get(name, 0)
    return rc


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL).  The kernels are launched on the CURRENT device's current
    stream, so a tensor living on another GPU is an error here rather than an illegal address later."""
    if t is None:
        return None
    if t.is_cuda:
        import torch
        cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise Vl3dError(f"tensor on cuda:{t.device.index} but the current device is cuda:{cur}: wrap the call in "
                            f"`with torch.cuda.device(tensor.device)`")
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
