"""Seeded synthetic workloads for tests and bench.py (BASELINE.json configs).

Everything is generated with numpy on the host from fixed seeds, so the CUDA path, the CPU oracle
and the reference arm of bench.py see bit-identical inputs.  Camera helpers restate
gluLookAt / gluPerspective (reference src/common/camera_utils.py:45-139) in vectorised numpy;
they are input generators, not part of the product path.
"""
import numpy as np

CUBE_VERTICES = np.array([[-1, -1, 1], [-1, -1, -1], [-1, 1, -1], [-1, 1, 1], [1, -1, 1],
                          [1, -1, -1], [1, 1, -1], [1, 1, 1]], np.float32)
CUBE_TRIANGLES = np.array([[0, 1, 2], [2, 3, 0], [3, 2, 6], [6, 7, 3], [7, 6, 5], [5, 4, 7],
                           [4, 5, 1], [1, 0, 4], [5, 6, 2], [2, 1, 5], [7, 4, 0], [0, 3, 7]], np.int32)


def perspective(aspect, fov_y_deg, near, far):
    """[4,4] OpenGL projection (camera_utils.py:99-139)."""
    f = 1.0 / np.tan(np.deg2rad(fov_y_deg) / 2.0)
    m = np.zeros((4, 4), np.float64)
    m[0, 0] = f / aspect
    m[1, 1] = f
    m[2, 2] = (near + far) / (near - far)
    m[2, 3] = 2.0 * near * far / (near - far)
    m[3, 2] = -1.0
    return m


def look_at(eye, center, up):
    """[N,4,4] world-to-eye matrices (camera_utils.py:45-96)."""
    eye, center, up = (np.asarray(a, np.float64).reshape(-1, 3) for a in (eye, center, up))
    fwd = center - eye
    fwd /= np.linalg.norm(fwd, axis=1, keepdims=True)
    side = np.cross(fwd, up)
    side /= np.linalg.norm(side, axis=1, keepdims=True)
    true_up = np.cross(side, fwd)
    rot = np.stack([side, true_up, -fwd], 1)                      # [N,3,3]
    m = np.tile(np.eye(4), (eye.shape[0], 1, 1))
    m[:, :3, :3] = rot
    m[:, :3, 3] = -np.einsum("nij,nj->ni", rot, eye)
    return m


def fibonacci_sphere(n, radius=1.0):
    i = np.arange(n) + 0.5
    phi = np.arccos(1.0 - 2.0 * i / n)
    theta = np.pi * (1.0 + 5.0 ** 0.5) * i
    return radius * np.stack([np.cos(theta) * np.sin(phi), np.cos(phi), np.sin(theta) * np.sin(phi)], 1)


def orbit_cameras(n, radius=3.0, fov_y=40.0, near=0.01, far=10.0, aspect=1.0):
    """n model-view-projection matrices [n,4,4] float32 looking at the origin from a Fibonacci sphere."""
    eye = fibonacci_sphere(n, radius)
    up = np.tile(np.array([[0.0, 1.0, 0.0]]), (n, 1))
    parallel = np.abs(eye[:, 1]) > 0.999 * radius
    up[parallel] = np.array([1.0, 0.0, 0.0])
    mv = look_at(eye, np.zeros((n, 3)), up)
    return (perspective(aspect, fov_y, near, far)[None] @ mv).astype(np.float32)


def uv_sphere(n_lon, n_rings, radius=1.0):
    """Closed UV sphere with proper longitude wrap: V = 2 + n_lon*n_rings, T = 2*n_lon*n_rings.

    (159, 158) -> T = 50 244, V = 25 124;  (224, 223) -> T = 99 904;  (708, 707) -> T = 1 001 112.
    """
    lat = np.pi * (np.arange(1, n_rings + 1) / (n_rings + 1))           # interior rings, pole excluded
    lon = 2.0 * np.pi * np.arange(n_lon) / n_lon
    st, ct = np.sin(lat)[:, None], np.cos(lat)[:, None]
    ring = np.stack([st * np.cos(lon)[None], np.broadcast_to(ct, (n_rings, n_lon)), st * np.sin(lon)[None]], -1)
    verts = np.concatenate([[[0.0, 1.0, 0.0]], ring.reshape(-1, 3), [[0.0, -1.0, 0.0]]], 0) * radius
    idx = 1 + np.arange(n_rings * n_lon).reshape(n_rings, n_lon)
    nxt = np.roll(idx, -1, axis=1)
    top = np.stack([np.zeros(n_lon, np.int64), nxt[0], idx[0]], 1)
    south = verts.shape[0] - 1
    bottom = np.stack([np.full(n_lon, south), idx[-1], nxt[-1]], 1)
    a, b, c, d = idx[:-1], nxt[:-1], idx[1:], nxt[1:]
    quads = np.concatenate([np.stack([a, b, d], -1).reshape(-1, 3), np.stack([a, d, c], -1).reshape(-1, 3)], 0)
    tris = np.concatenate([top, quads, bottom], 0).astype(np.int32)
    return verts.astype(np.float32), tris


def transform(mvp, world_vertices):
    """[B,4,4] x [V,3] or [B,V,3] -> clip-space [B,V,4] float32 (fp32 matmul like camera_utils.py:166-170)."""
    mvp = np.asarray(mvp, np.float32)
    w = np.asarray(world_vertices, np.float32)
    if w.ndim == 2:
        w = np.broadcast_to(w[None], (mvp.shape[0],) + w.shape)
    hom = np.concatenate([w, np.ones(w.shape[:2] + (1,), np.float32)], 2)
    return np.einsum("bvk,bjk->bvj", hom, mvp).astype(np.float32)


def euler_matrices(angles):
    """[N,3] (x, y, z angles in radians) -> [N,3,3], composition of camera_utils.py:10-42."""
    a = np.asarray(angles, np.float64).reshape(-1, 3)
    s, c = np.sin(a), np.cos(a)
    sx, sy, sz, cx, cy, cz = s[:, 0], s[:, 1], s[:, 2], c[:, 0], c[:, 1], c[:, 2]
    m = np.empty((a.shape[0], 3, 3))
    m[:, 0, 0] = cz * cy; m[:, 0, 1] = cz * sy * sx - cx * sz; m[:, 0, 2] = sz * sx + cz * cx * sy
    m[:, 1, 0] = cy * sz; m[:, 1, 1] = cz * cx + sz * sy * sx; m[:, 1, 2] = cx * sz * sy - cz * sx
    m[:, 2, 0] = -sy; m[:, 2, 1] = cy * sx; m[:, 2, 2] = cy * cx
    return m


def sphere_views(n_lon, n_rings, batch, size, attributes=9, seed=0, radius=3.0):
    """Configs c2 / c3 / c4: one UV sphere seen from `batch` orbit cameras.

    Returns dict(clip_vertices [B,V,4], attributes [B,V,A], triangles [T,3], background [A],
    width, height, world_vertices [V,3], camera_matrices [B,4,4]).
    """
    verts, tris = uv_sphere(n_lon, n_rings)
    mvp = orbit_cameras(batch, radius=radius)
    rng = np.random.default_rng(seed)
    attrs = rng.random((batch, verts.shape[0], attributes), dtype=np.float32)
    return dict(clip_vertices=transform(mvp, verts), attributes=attrs, triangles=tris,
                background=-np.ones(attributes, np.float32), width=size, height=size,
                world_vertices=verts, camera_matrices=mvp)


def cube_test_scene(width=640, height=480):
    """Config c1: the geometry of mesh_renderer_test.py:30-57 (two Euler-rotated cubes, eye z=6,
    fov 40, near 0.01, far 10), attributes = [normals, world positions, ones] (A = 9,
    render.py:181), background -1 (render.py:197)."""
    rot = euler_matrices([[-20.0, 0.0, 60.0], [45.0, 60.0, 0.0]]).astype(np.float32)
    world = np.einsum("vk,bjk->bvj", CUBE_VERTICES, rot).astype(np.float32)
    normals = CUBE_VERTICES / np.linalg.norm(CUBE_VERTICES, axis=1, keepdims=True)
    normals_w = np.einsum("vk,bjk->bvj", normals.astype(np.float32), rot).astype(np.float32)
    mv = look_at(np.tile([[0.0, 0.0, 6.0]], (2, 1)), np.zeros((2, 3)), np.tile([[0.0, 1.0, 0.0]], (2, 1)))
    mvp = (perspective(width / height, 40.0, 0.01, 10.0)[None] @ mv).astype(np.float32)
    attrs = np.concatenate([normals_w, world, np.ones_like(world)], 2).astype(np.float32)
    return dict(clip_vertices=transform(mvp, world), attributes=attrs, triangles=CUBE_TRIANGLES.copy(),
                background=-np.ones(9, np.float32), width=width, height=height,
                world_vertices=world, camera_matrices=mvp)


def occlusion_soup(batch, size, n_triangles=2000, scale=0.35, attributes=9, seed=5):
    """Config c5: `n_triangles` large independent triangles per image (V = 3T), random winding,
    per-vertex w in [0.5, 2] (the perspective trick of rasterize_triangles_test.py:52-53); `scale`
    is calibrated so that the mean depth complexity (inside-test passes per pixel) is ~50 at
    2000 triangles."""
    clips = []
    for b in range(batch):
        rng = np.random.default_rng(seed + b)
        centre = rng.uniform(-0.8, 0.8, (n_triangles, 1, 2))
        xy = centre + scale * rng.standard_normal((n_triangles, 3, 2))
        z = rng.uniform(-0.9, 0.9, (n_triangles, 3, 1))
        w = rng.uniform(0.5, 2.0, (n_triangles, 3, 1))
        v = np.concatenate([xy, z, np.ones_like(z)], 2) * w
        flip = rng.random(n_triangles) < 0.5
        v[flip] = v[flip][:, ::-1]
        clips.append(v.reshape(3 * n_triangles, 4))
    clip = np.stack(clips).astype(np.float32)
    tris = np.arange(3 * n_triangles, dtype=np.int32).reshape(n_triangles, 3)
    rng = np.random.default_rng(seed + 1000)
    attrs = rng.random((batch, 3 * n_triangles, attributes), dtype=np.float32)
    return dict(clip_vertices=clip, attributes=attrs, triangles=tris,
                background=-np.ones(attributes, np.float32), width=size, height=size)


def upstream_gradient(shape, seed=1):
    return np.random.default_rng(seed).standard_normal(tuple(shape), dtype=np.float32)
