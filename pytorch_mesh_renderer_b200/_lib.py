"""ctypes binding of libpmr_b200.so (include/pmr_b200.h).  Thin by design: tensors cross as raw
device pointers plus sizes, torch supplies memory and the current stream.

There is no fallback of any kind: if the shared library is missing or no CUDA device is
present, the first call raises.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpmr_b200.so")

BACKWARD_ATOMIC = 0
BACKWARD_ORDERED = 1

_lib = None
_contexts = {}
_lock = threading.Lock()

_vp = ctypes.c_void_p
_i = ctypes.c_int

_PROTOTYPES = {
    "pmr_version": (ctypes.c_int, []),
    "pmr_create": (ctypes.c_int, [_i, ctypes.POINTER(_vp)]),
    "pmr_destroy": (None, [_vp]),
    "pmr_last_error": (ctypes.c_char_p, [_vp]),
    "pmr_launch_count": (ctypes.c_longlong, [_vp]),
    "pmr_last_large_triangles": (ctypes.c_longlong, [_vp]),
    "pmr_set_small_mesh_threshold": (ctypes.c_int, [_vp, _i]),
    "pmr_enable_stage_timing": (ctypes.c_int, [_vp, _i]),
    "pmr_read_stage_timing": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_longlong), _i]),
    "pmr_rasterize_forward": (ctypes.c_int, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "pmr_rasterize_backward": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "pmr_interpolate_forward": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "pmr_rasterize_interpolate_forward": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i,
                                                         _vp, _vp, _vp, _vp, _vp]),
    "pmr_rasterize_interpolate_backward": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i,
                                                          _vp, _vp, _i, _vp]),
    "pmr_transform_forward": (ctypes.c_int, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "pmr_transform_backward": (ctypes.c_int, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "pmr_peer_exchange_bytes": (ctypes.c_size_t, [ctypes.c_longlong, _i]),
    "pmr_peer_alloc": (ctypes.c_int, [_vp, ctypes.c_size_t, ctypes.POINTER(_vp), _vp]),
    "pmr_peer_open": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(_vp)]),
    "pmr_peer_close": (ctypes.c_int, [_vp, _vp]),
    "pmr_peer_free": (ctypes.c_int, [_vp, _vp]),
    "pmr_peer_status": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(_i)]),
    "pmr_transform_backward_exchange": (ctypes.c_int, [_vp, _vp, _vp, _i, _i, ctypes.POINTER(_vp), _i, _i,
                                                       ctypes.c_longlong, _vp, _vp]),
    "pmr_vertex_incidence": (ctypes.c_int, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    "pmr_vertex_normals_forward": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "pmr_vertex_normals_backward": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "pmr_shade_diffuse_forward": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "pmr_shade_diffuse_backward": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "pmr_shade_phong_forward": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "pmr_shade_phong_backward": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i,
                                                _vp, _vp, _vp]),
    "pmr_render_diffuse_forward": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i,
                                                  _vp, _vp, _vp, _vp, _vp]),
    "pmr_render_diffuse_backward": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                                   _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "pmr_rasterize_clip_space_host": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i,
                                                     _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
}

EXPORTED_SYMBOLS = tuple(_PROTOTYPES)


def load():
    """dlopen the library and set prototypes.  Needs no GPU (symbol checks run on CPU boxes)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "pytorch_mesh_renderer_b200: %s is missing -- build it with "
                "`python -m pytorch_mesh_renderer_b200.build` (there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class PmrError(RuntimeError):
    pass


def context(device_index):
    """One library context (device workspace) per CUDA device, created lazily."""
    with _lock:
        ctx = _contexts.get(device_index)
        if ctx is None:
            lib = load()
            if not torch.cuda.is_available():
                raise RuntimeError("pytorch_mesh_renderer_b200 needs a CUDA device; there is no CPU fallback")
            handle = _vp()
            # pmr_create (like every entry point) makes `device_index` the calling thread's current CUDA
            # device; the guard restores the caller's device afterwards.
            with torch.cuda.device(int(device_index)):
                rc = lib.pmr_create(int(device_index), ctypes.byref(handle))
            if rc != 0:
                raise PmrError("pmr_create(device=%d) failed with code %d" % (device_index, rc))
            ctx = handle
            _contexts[device_index] = ctx
        return ctx


def check(ctx, rc):
    if rc == 0:
        return
    msg = load().pmr_last_error(ctx)
    msg = msg.decode() if msg else "error"
    if rc == -1:
        raise ValueError(msg)          # same exception type as rasterize.py:98-103
    raise PmrError("libpmr_b200: %s (code %d)" % (msg, rc))


def ptr(t):
    return _vp(t.data_ptr()) if t is not None else _vp(0)


def stream_ptr(device):
    return _vp(torch.cuda.current_stream(device).cuda_stream)


def launch_count(device_index=None):
    lib = load()
    if device_index is None:
        return sum(lib.pmr_launch_count(c) for c in _contexts.values())
    c = _contexts.get(device_index)
    return lib.pmr_launch_count(c) if c is not None else 0


STAGES = ("bin", "raster", "backward", "interp", "scatter", "resolve", "shade")


def enable_stage_timing(device_index, on=True):
    check(context(device_index), load().pmr_enable_stage_timing(context(device_index), int(bool(on))))


def read_stage_timing(device_index, reset=True):
    """{stage: (total_ms, intervals)} accumulated since the last reset."""
    ms = (ctypes.c_double * len(STAGES))()
    n = (ctypes.c_longlong * len(STAGES))()
    ctx = context(device_index)
    check(ctx, load().pmr_read_stage_timing(ctx, ms, n, int(bool(reset))))
    return {name: (ms[k], n[k]) for k, name in enumerate(STAGES)}
