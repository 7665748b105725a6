"""Deterministic synthetic scenes and ray sets.  numpy only; used by tests/ and
bench.py to feed both the CUDA library and the CPU oracle with identical inputs.

Meshes are triangle soups: float32 arrays of shape (n_faces, 9) = (v0, v1, v2) per face, the
form the C ABI takes (rc_push).  Ray sets are RTRay records (src/rt_transport.jl:10-19).
"""
from __future__ import annotations

import numpy as np

RAY_DTYPE = np.dtype([("o", "<f4", 3), ("t_min", "<f4"), ("d", "<f4", 3), ("t_max", "<f4")])
HIT_DTYPE = np.dtype(
    [
        ("hit", "<u4"),
        ("t", "<f4"),
        ("primitive_id", "<u4"),
        ("instance_custom_index", "<u4"),
        ("bary_u", "<f4"),
        ("bary_v", "<f4"),
        ("instance_id", "<u4"),
        ("meta", "<u4"),
    ]
)

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def rng_uniform(seed: int, index, dim: int) -> np.ndarray:
    """Counter-based uniform in [0,1) with 24 random bits: splitmix64 finaliser of
    seed + GOLDEN*(4*index + dim + 1).  Same specification as the CUDA library's device RNG
    (csrc/rc_common.cuh rc_rng_uniform) and the oracle's orc_rng_uniform."""
    idx = np.asarray(index, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + np.uint64(0x9E3779B97F4A7C15) * (idx * np.uint64(4) + np.uint64(dim + 1))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)


# ------------------------------------------------------------------------------- meshes
def _grid_sphere(n: int, center, radius_fn) -> np.ndarray:
    theta = np.pi * np.arange(n, dtype=np.float64) / (n - 1)
    phi = 2.0 * np.pi * np.arange(n, dtype=np.float64) / (n - 1)
    T, P = np.meshgrid(theta, phi, indexing="ij")
    R = radius_fn(T, P)
    pts = np.stack([R * np.sin(T) * np.cos(P), R * np.sin(T) * np.sin(P), R * np.cos(T)], axis=-1)
    pts = (pts + np.asarray(center, np.float64)).astype(np.float32)
    a = pts[:-1, :-1]  # (i, j)
    b = pts[1:, :-1]  # (i+1, j)
    c = pts[1:, 1:]  # (i+1, j+1)
    d = pts[:-1, 1:]  # (i, j+1)
    t1 = np.concatenate([a, b, c], axis=-1)  # outward winding
    t2 = np.concatenate([a, c, d], axis=-1)
    tris = np.stack([t1, t2], axis=2).reshape(-1, 9)
    return np.ascontiguousarray(tris, dtype=np.float32)


def uv_sphere(n: int, center=(0.0, 0.0, 0.0), r: float = 1.0) -> np.ndarray:
    """n x n lat/long tessellation (topology of GeometryBasics' Tesselation(Sphere, n), not bit-identical):
    2(n-1)^2 faces including the degenerate pole faces the reference's filter drops."""
    return _grid_sphere(n, center, lambda T, P: np.full_like(T, r))


def bumpy_sphere(n: int, center=(0.0, 0.0, 0.0), r0: float = 1.0) -> np.ndarray:
    """radius r0*(1 + 0.15 sin(7 theta) sin(5 phi)) -- non-convex, so interior rays see several candidates."""
    return _grid_sphere(n, center, lambda T, P: r0 * (1.0 + 0.15 * np.sin(7 * T) * np.sin(5 * P)))


def is_degenerate(tris: np.ndarray) -> np.ndarray:
    """is_degenerate (src/triangle_mesh.jl:14-17) for a float32 (n, 9) soup, evaluated in Float32 like the reference:
    c = (v3 - v1) x (v2 - v1), degenerate iff dot(c, c) == 0 (`≈ 0f0` on Float32 is an exact test)."""
    v = np.asarray(tris, np.float32).reshape(-1, 3, 3)
    a, b = v[:, 2] - v[:, 0], v[:, 1] - v[:, 0]
    cx = a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1]
    cy = a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2]
    cz = a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]
    return ((cx * cx + cy * cy) + cz * cz) == 0


def box_mesh(lo=(-0.5, -0.5, -0.5), hi=(0.5, 0.5, 0.5)) -> np.ndarray:
    """12-triangle axis-aligned box (stress_box analogue, test/test_tlas_stress.jl:40), outward winding."""
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    p = np.array(
        [[x0, y0, z0], [x1, y0, z0], [x1, y1, z0], [x0, y1, z0], [x0, y0, z1], [x1, y0, z1], [x1, y1, z1], [x0, y1, z1]], np.float32
    )
    quads = [(0, 3, 2, 1), (4, 5, 6, 7), (0, 1, 5, 4), (2, 3, 7, 6), (1, 2, 6, 5), (3, 0, 4, 7)]
    tris = []
    for a, b, c, d in quads:
        tris.append(np.concatenate([p[a], p[b], p[c]]))
        tris.append(np.concatenate([p[a], p[c], p[d]]))
    return np.ascontiguousarray(np.stack(tris), np.float32)


def quad_mesh(z: float = 0.0, half: float = 1.0) -> np.ndarray:
    """two-triangle square in the plane z (make_test_mesh analogue, test/test_instanced_bvh.jl:412-415)."""
    p = np.array([[-half, -half, z], [half, -half, z], [half, half, z], [-half, half, z]], np.float32)
    return np.ascontiguousarray(np.stack([np.concatenate([p[0], p[1], p[2]]), np.concatenate([p[0], p[2], p[3]])]), np.float32)


# ------------------------------------------------------------------------------- transforms (Mat3x4f rows)
def identity3x4() -> np.ndarray:
    return np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)


def translation3x4(t) -> np.ndarray:
    m = identity3x4()
    m[3], m[7], m[11] = t
    return m


def trs3x4(t, quat, s) -> np.ndarray:
    """T*R*S as Vulkan row-major 3x4; quat = (w,x,y,z) unit."""
    w, x, y, z = [float(q) for q in quat]
    R = np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
        ]
    )
    m = np.zeros((3, 4))
    m[:, :3] = R * float(s)
    m[:, 3] = t
    return m.astype(np.float32).reshape(12)


def random_trs(n: int, seed: int, extent: float = 40.0, smin: float = 0.5, smax: float = 1.5) -> np.ndarray:
    """n random T*R*S transforms: T~U[-extent,extent]^3, R uniform unit quaternion, S~U[smin,smax] (C3)."""
    rs = np.random.RandomState(seed)
    t = rs.uniform(-extent, extent, (n, 3))
    q = rs.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    s = rs.uniform(smin, smax, n)
    return np.stack([trs3x4(t[i], q[i], s[i]) for i in range(n)]).astype(np.float32)


# ------------------------------------------------------------------------------- ray sets
def make_rays(o, d, t_min=0.0, t_max=np.inf) -> np.ndarray:
    o = np.asarray(o, np.float32).reshape(-1, 3)
    d = np.broadcast_to(np.asarray(d, np.float32).reshape(-1, 3), o.shape)
    r = np.zeros(len(o), RAY_DTYPE)
    r["o"], r["d"], r["t_min"], r["t_max"] = o, d, t_min, t_max
    return r


def pinhole_rays(width: int, height: int, camera_pos=(0.0, 0.0, -3.0), fov_deg: float = 45.0) -> np.ndarray:
    """camera_ray of docs/src/raytracing-core.jl:12-17 looking down +z, row-major pixel order."""
    focal = np.float32(1.0 / np.tan(np.radians(fov_deg) / 2.0))
    aspect = np.float32(width / height)
    x = np.arange(1, width + 1, dtype=np.float32)
    y = np.arange(1, height + 1, dtype=np.float32)
    ndc_x = (np.float32(2) * (x - np.float32(0.5)) / np.float32(width) - np.float32(1)) * aspect
    ndc_y = np.float32(1) - np.float32(2) * (y - np.float32(0.5)) / np.float32(height)
    X, Y = np.meshgrid(ndc_x, ndc_y, indexing="xy")
    d = np.stack([X, Y, np.full_like(X, focal)], axis=-1).reshape(-1, 3).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    return make_rays(np.broadcast_to(np.asarray(camera_pos, np.float32), d.shape), d)


def uniform_sphere_dirs(n: int, seed: int, first_index: int = 0) -> np.ndarray:
    idx = np.arange(first_index, first_index + n, dtype=np.uint64)
    u1, u2 = rng_uniform(seed, idx, 0), rng_uniform(seed, idx, 1)
    z = 1.0 - 2.0 * u1.astype(np.float64)
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    phi = 2.0 * np.pi * u2.astype(np.float64)
    return np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=-1).astype(np.float32)


def interior_rays(n: int, seed: int, radius: float = 0.8, first_index: int = 0) -> np.ndarray:
    """origins uniform in the ball |x| < radius, directions uniform on S^2 (set C)."""
    idx = np.arange(first_index, first_index + n, dtype=np.uint64)
    d0 = uniform_sphere_dirs(n, seed ^ 0xA5A5, first_index)
    rr = radius * np.cbrt(rng_uniform(seed, idx, 2).astype(np.float64))
    o = (d0.astype(np.float64) * rr[:, None]).astype(np.float32)
    d = uniform_sphere_dirs(n, seed, first_index)
    return make_rays(o, d)


def box_rays(n: int, seed: int, half: float = 44.0, first_index: int = 0) -> np.ndarray:
    """origins ~U[-half,half]^3, directions uniform on S^2 (C3 ray set)."""
    idx = np.arange(first_index, first_index + n, dtype=np.uint64)
    o = np.stack([(rng_uniform(seed ^ 0x1234, idx, k).astype(np.float64) * 2 - 1) * half for k in range(3)], axis=-1).astype(np.float32)
    d = uniform_sphere_dirs(n, seed, first_index)
    return make_rays(o, d)


def bounce_rays(n: int, primary: np.ndarray, hits: np.ndarray, normals: np.ndarray, seed: int, offset: float = 1e-3) -> np.ndarray:
    """Diffuse-bounce rays (set B): ray k starts at primary hit (k mod n_hits) + offset*n_geo and leaves in a
    direction uniform over the hemisphere about n_geo (n_geo flipped against the incoming ray)."""
    sel = np.nonzero(hits["hit"])[0]
    assert len(sel) > 0
    k = np.arange(n, dtype=np.int64)
    src = sel[k % len(sel)]
    pr = primary[src]
    p = pr["o"].astype(np.float64) + pr["d"].astype(np.float64) * hits["t"][src].astype(np.float64)[:, None]
    ng = normals[src].astype(np.float64)
    flip = np.sum(ng * pr["d"], axis=1) > 0
    ng[flip] *= -1
    d = uniform_sphere_dirs(n, seed).astype(np.float64)
    neg = np.sum(d * ng, axis=1) < 0
    d[neg] *= -1
    return make_rays((p + offset * ng).astype(np.float32), d.astype(np.float32))


def geometric_normals(tris: np.ndarray) -> np.ndarray:
    v = tris.reshape(-1, 3, 3).astype(np.float64)
    n = np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0])
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    ln[ln == 0] = 1
    return (n / ln).astype(np.float32)


def viewfactor_scene(n: int = 72):
    """Five bumpy spheres (C4): list of (verts, center)."""
    centers = [(0.0, 0.0, 0.0), (2.6, 0.0, 0.0), (-2.6, 0.0, 0.0), (0.0, 2.6, 0.0), (0.0, 0.0, 2.6)]
    radii = [1.0, 0.8, 0.8, 0.6, 0.6]
    return [bumpy_sphere(n, c, r) for c, r in zip(centers, radii)]
