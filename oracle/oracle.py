"""ctypes front-end of oracle/raster_oracle.c (plain-C restatement of the reference path).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  All arrays are numpy, C-contiguous,
float32 / int32.  Every function cites the reference lines it restates in raster_oracle.c.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libraster_oracle.so")
_SRC = os.path.join(_HERE, "raster_oracle.c")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-ffp-contract=off",
             "-fno-fast-math", "-o", _SO, _SRC, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        L.pmr_oracle_forward.argtypes = [_f32p, _i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         _i32p, _f32p, _f32p, _i64p]
        L.pmr_oracle_forward.restype = None
        L.pmr_oracle_backward.argtypes = [_f32p, ctypes.c_int64, _f32p, _i32p, _i32p, _f32p,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p]
        L.pmr_oracle_backward.restype = None
        L.pmr_oracle_backward_f64acc.argtypes = [_f32p, ctypes.c_int64, _f32p, _i32p, _i32p, _f32p,
                                                 ctypes.c_int, ctypes.c_int, ctypes.c_int, _f64p]
        L.pmr_oracle_backward_f64acc.restype = None
        L.pmr_oracle_interp_forward.argtypes = [_f32p, _i32p, _i32p, _f32p, _f32p,
                                                ctypes.c_int, ctypes.c_int64, _f32p]
        L.pmr_oracle_interp_forward.restype = None
        L.pmr_oracle_interp_backward.argtypes = [_f32p, _f32p, _i32p, _i32p, _f32p, _f32p,
                                                 ctypes.c_int, ctypes.c_int, ctypes.c_int64,
                                                 _f32p, _f32p]
        L.pmr_oracle_interp_backward.restype = None
        L.pmr_oracle_vertex_normals.argtypes = [_f32p, _i32p, ctypes.c_int, ctypes.c_int, _f32p, _f32p]
        L.pmr_oracle_vertex_normals.restype = None
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, t):
    return a.ctypes.data_as(t)


def forward(vertices, triangles, image_width, image_height, return_counters=False):
    """One image.  vertices [V,4], triangles [T,3] -> ids [H,W] i32, bary [H,W,3], z [H,W]."""
    v, t = _f32(vertices), _i32(triangles)
    assert v.ndim == 2 and v.shape[1] == 4 and t.ndim == 2 and t.shape[1] == 3
    W, H = int(image_width), int(image_height)
    ids = np.empty((H, W), np.int32)
    bary = np.empty((H, W, 3), np.float32)
    z = np.empty((H, W), np.float32)
    counters = np.zeros(3, np.int64)
    lib().pmr_oracle_forward(_p(v, _f32p), _p(t, _i32p), t.shape[0], W, H,
                             _p(ids, _i32p), _p(bary, _f32p), _p(z, _f32p), _p(counters, _i64p))
    if return_counters:
        return ids, bary, z, counters
    return ids, bary, z


def backward(df_dbary, vertices, triangles, ids, bary):
    """One image.  -> df_dvertices [V,4] (columns x, y, w; z column zero)."""
    g, v, t, i, b = _f32(df_dbary), _f32(vertices), _i32(triangles), _i32(ids), _f32(bary)
    H, W = i.shape
    out = np.empty((v.shape[0], 4), np.float32)
    lib().pmr_oracle_backward(_p(g, _f32p), 3, _p(v, _f32p), _p(t, _i32p), _p(i, _i32p),
                              _p(b, _f32p), v.shape[0], W, H, _p(out, _f32p))
    return out


def backward_f64acc(df_dbary, vertices, triangles, ids, bary):
    """Same fp32 per-pixel terms summed in double (SURVEY.md F5 yardstick)."""
    g, v, t, i, b = _f32(df_dbary), _f32(vertices), _i32(triangles), _i32(ids), _f32(bary)
    H, W = i.shape
    out = np.empty((v.shape[0], 4), np.float64)
    lib().pmr_oracle_backward_f64acc(_p(g, _f32p), 3, _p(v, _f32p), _p(t, _i32p), _p(i, _i32p),
                                     _p(b, _f32p), v.shape[0], W, H, _p(out, _f64p))
    return out


def interp_forward(attributes, triangles, ids, bary, background):
    """One image.  attributes [V,A] -> out [H,W,A]  (rast.py:118-150)."""
    a, t, i, b, bg = _f32(attributes), _i32(triangles), _i32(ids), _f32(bary), _f32(background)
    H, W = i.shape
    A = a.shape[1]
    out = np.empty((H, W, A), np.float32)
    lib().pmr_oracle_interp_forward(_p(a, _f32p), _p(t, _i32p), _p(i, _i32p), _p(b, _f32p),
                                    _p(bg, _f32p), A, H * W, _p(out, _f32p))
    return out


def interp_backward(grad_out, attributes, triangles, ids, bary, background):
    """One image.  -> (d_attributes [V,A], d_bary [H,W,3])  (autograd of rast.py:118-150)."""
    g, a, t = _f32(grad_out), _f32(attributes), _i32(triangles)
    i, b, bg = _i32(ids), _f32(bary), _f32(background)
    H, W = i.shape
    V, A = a.shape
    dattr = np.empty((V, A), np.float32)
    dbary = np.empty((H, W, 3), np.float32)
    lib().pmr_oracle_interp_backward(_p(g, _f32p), _p(a, _f32p), _p(t, _i32p), _p(i, _i32p),
                                     _p(b, _f32p), _p(bg, _f32p), V, A, H * W,
                                     _p(dattr, _f32p), _p(dbary, _f32p))
    return dattr, dbary


def interp_backward_attributes_f64acc(grad_out, triangles, ids, bary, vertex_count):
    """d_attributes with the same fp32 products g_a*alpha*b_k but summed in double (yardstick for
    the atomic mode)."""
    g, t, i, b = _f32(grad_out), _i32(triangles), _i32(ids).reshape(-1), _f32(bary).reshape(-1, 3)
    A = g.shape[-1]
    g = g.reshape(-1, A)
    s = (np.float32(2.0) * b[:, 0] + np.float32(2.0) * b[:, 1]) + np.float32(2.0) * b[:, 2]
    alpha = np.clip(s, np.float32(0.0), np.float32(1.0)).astype(np.float32)
    d_img = (g * alpha[:, None]).astype(np.float32)
    out = np.zeros((vertex_count, A), np.float64)
    if t.shape[0] == 0:
        return out
    for k in range(3):
        np.add.at(out, t[i, k], (d_img * b[:, k:k + 1]).astype(np.float32).astype(np.float64))
    return out


def atomic_mode_bound(ref32, ref64, slack=8.0):
    """Per-entry bound for a result that sums the reference's fp32 terms in another order: the
    north-star tolerance around the exactly summed value plus `slack` times the reference's own
    worst distance from it (the reference is itself only one particular fp32 order)."""
    ref_err = np.abs(np.asarray(ref32, np.float64) - ref64).max() if ref64.size else 0.0
    return 1e-6 + 1e-5 * np.abs(ref64) + slack * ref_err


def tolerance_report(mine, ref32, ref64):
    """How a gradient computed in another summation order stands against the north-star tolerance
    1e-6 + 1e-5*|ref| itself: the fraction of entries outside it (a) against the reference's own fp32 result,
    (b) against the exactly summed fp32 terms, and -- the context for (b) -- the reference's own fraction
    against the same yardstick, with the largest absolute errors."""
    mine = np.asarray(mine, np.float64)
    ref32 = np.asarray(ref32, np.float64)
    ref64 = np.asarray(ref64, np.float64)
    n = max(mine.size, 1)
    outside = lambda a, b: float((np.abs(a - b) > 1e-6 + 1e-5 * np.abs(b)).sum()) / n
    biggest = lambda a, b: float(np.abs(a - b).max()) if a.size else 0.0
    return {"entries": int(mine.size),
            "outside_vs_reference_order": outside(mine, ref32),
            "outside_vs_f64_sum": outside(mine, ref64),
            "reference_outside_vs_f64_sum": outside(ref32, ref64),
            "max_abs_err_vs_f64_sum": biggest(mine, ref64),
            "reference_max_abs_err_vs_f64_sum": biggest(ref32, ref64),
            "max_abs_value": float(np.abs(ref64).max()) if ref64.size else 0.0}


def rasterize_clip_space(clip_space_vertices, attributes, triangles, image_width, image_height,
                         background_value, grad_out=None, f64_yardstick=False):
    """Batched restatement of rast.py:66-152 (a Python loop over images, like rast.py:112).

    Returns dict(out, ids, bary, z) and, when grad_out [B,H,W,A] is given, also
    d_vertices [B,V,4] and d_attributes [B,V,A] (plus their double-accumulated versions when
    f64_yardstick is set).
    """
    cv, at, tr = _f32(clip_space_vertices), _f32(attributes), _i32(triangles)
    bg = _f32(np.broadcast_to(np.asarray(background_value, np.float32), (at.shape[2],)))
    B = cv.shape[0]
    res = dict(out=[], ids=[], bary=[], z=[])
    if grad_out is not None:
        res.update(d_vertices=[], d_attributes=[])
        if f64_yardstick:
            res.update(d_vertices_f64=[], d_attributes_f64=[])
    for b in range(B):
        ids, bary, z = forward(cv[b], tr, image_width, image_height)
        res["ids"].append(ids)
        res["bary"].append(bary)
        res["z"].append(z)
        res["out"].append(interp_forward(at[b], tr, ids, bary, bg))
        if grad_out is not None:
            dattr, dbary = interp_backward(grad_out[b], at[b], tr, ids, bary, bg)
            res["d_attributes"].append(dattr)
            res["d_vertices"].append(backward(dbary, cv[b], tr, ids, bary))
            if f64_yardstick:
                res["d_vertices_f64"].append(backward_f64acc(dbary, cv[b], tr, ids, bary))
                res["d_attributes_f64"].append(
                    interp_backward_attributes_f64acc(grad_out[b], tr, ids, bary, cv.shape[1]))
    return {k: np.stack(v) for k, v in res.items()}


def vertex_normals(vertices, triangles, return_raw=False):
    """compute_vertex_normals (src/common/meshes.py:3-35), batched like the reference: vertices [B,V,3]."""
    v, t = _f32(vertices), _i32(triangles)
    assert v.ndim == 3 and v.shape[2] == 3
    normals, raw = np.empty_like(v), np.empty_like(v)
    for b in range(v.shape[0]):
        lib().pmr_oracle_vertex_normals(_p(v[b], _f32p), _p(t, _i32p), v.shape[1], t.shape[0],
                                        _p(raw[b], _f32p), _p(normals[b], _f32p))
    return (normals, raw) if return_raw else normals
