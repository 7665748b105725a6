"""Serialised geometry (`to_gpu(ArrayType, blas::BLAS)`, src/kernel-abstractions.jl:31-36): a geometry restored from
an exported blob gives byte-identical structures and byte-identical traces, without running the builder; damaged blobs are refused."""
import ctypes as C

import numpy as np
import pytest

import raycore_b200 as rc
from raycore_b200 import _lib as L
from raycore_b200 import tlas as T
from raycore_b200 import workloads as W
from raycore_b200._lib import RAY_DTYPE

pytestmark = pytest.mark.gpu
F = np.float32
M64 = (1 << 64) - 1


def blob_hash(payload: bytes) -> int:
    """blob_hash of csrc/rc_build.cu restated (word-wise multiply-xorshift over the payload)."""
    h = 0x9E3779B97F4A7C15 ^ len(payload)
    nw = len(payload) // 8
    for w in np.frombuffer(payload, "<u8", count=nw).tolist():
        h = ((h ^ w) * 0xD6E8FEB86659FD93) & M64
        h ^= h >> 32
    for b in payload[8 * nw:]:
        h = ((h ^ b) * 0x100000001B3) & M64
    return h


def _scene():
    sphere = W.bumpy_sphere(40)
    box = np.concatenate([W.box_mesh()[:5], np.zeros((1, 9), F), W.box_mesh()[5:]])  # a degenerate face in the submitted soup
    xs, xb = W.random_trs(30, seed=3, extent=6.0), W.random_trs(12, seed=4, extent=6.0)
    ids = np.arange(100, 130, dtype=np.uint32)
    return sphere, box, xs, xb, ids


def test_export_import_round_trip_is_byte_identical():
    sphere, box, xs, xb, ids = _scene()
    a = rc.TLAS(keep_bvh2=True)
    hs_ = a.push(sphere, list(xs), instance_ids=ids)
    hb = a.push(box, list(xb))
    rng = np.random.default_rng(5)
    nb = rng.standard_normal((len(box), 9)).astype(F)
    a.set_normals(hb, nb)
    a.sync()
    rays = np.concatenate([W.box_rays(150_000, 1, half=8.0), W.interior_rays(150_000, 2, radius=7.0)])
    ha_c, ha_a = a.trace_closest(rays), a.trace_any(rays)
    ha_r = a.trace_closest(rays, reference_order=True)
    assert 0.05 < ha_c["hit"].mean() < 0.95

    blob_s, blob_b = a.export_geometry(hs_), a.export_geometry(hb)
    hd_s, hd_b = T.blob_header(blob_s), T.blob_header(blob_b)
    assert hd_s["magic"] == b"RCBLAS\x00\x02" and hd_s["total_bytes"] == blob_s.nbytes and hd_s["has_normals"] == 0
    assert hd_b["n"] == len(W.box_mesh()) and hd_b["n_faces_in"] == len(box) and hd_b["has_normals"] == 1
    assert blob_hash(blob_s[128:].tobytes()) == int(hd_s["payload_hash"])
    assert np.array_equal(a.export_geometry(hs_), blob_s), "export is not deterministic"
    # the blob keeps the submitted soup (minus dropped faces): same vertices, face for face
    faces = T.blob_faces(blob_b)
    keep = np.array([i for i in range(len(box)) if i != 5])
    assert np.array_equal(faces[keep], box[keep]) and not faces[5].any()

    b = rc.TLAS(keep_bvh2=True)
    gs = b.push_exported(blob_s.tobytes(), list(xs), instance_ids=ids)  # bytes and arrays are both accepted
    gb = b.push_exported(blob_b, list(xb))
    assert b.n_geometries() == 2 and b.n_instances() == 42 and b.dirty
    b.sync()
    assert b.last_sync_action == rc.RC_SYNC_REBUILD
    for k in (1, 2):
        assert a.read_blas_nodes(k).tobytes() == b.read_blas_nodes(k).tobytes()
        assert np.array_equal(a.read_blas_order(k), b.read_blas_order(k))
        assert np.array_equal(a.read_blas_faces(k), b.read_blas_faces(k))
    assert a.read_tlas_nodes().tobytes() == b.read_tlas_nodes().tobytes()
    assert np.array_equal(a.get_instances(hs_), b.get_instances(gs))
    assert b.trace_closest(rays).tobytes() == ha_c.tobytes()
    assert b.trace_any(rays).tobytes() == ha_a.tobytes()
    assert b.trace_closest(rays, reference_order=True).tobytes() == ha_r.tobytes()
    # a blob of the restored geometry is the blob it came from
    assert np.array_equal(b.export_geometry(gs), blob_s) and np.array_equal(b.export_geometry(gb), blob_b)

    # shading-side data travels too: the shadow stage (interpolated normals) gives the same visibility
    lights = np.array([[0, 9, 0], [5, -3, 2]], F)
    vis = []
    for t in (a, b):
        q = t.queue(RAY_DTYPE, len(rays)).upload(rays)
        vis.append(t.shadow_visibility(q, t.intersect_rays(q), lights).download())
    assert vis[0].tobytes() == vis[1].tobytes() and 0 < vis[0].sum() < vis[0].size

    # per-ray API on a restored geometry materialises the Triangle from the blob's soup
    k = int(np.nonzero(ha_c["hit"])[0][0])
    ray = rc.Ray(rays[k]["o"], rays[k]["d"], t_min=float(rays[k]["t_min"]), t_max=float(rays[k]["t_max"]))
    ta, tb = a.closest_hit(ray), b.closest_hit(ray)
    assert ta[0] and tb[0] and np.array_equal(ta[1].vertices, tb[1].vertices) and ta[1].metadata == tb[1].metadata
    assert ta[2] == tb[2] and np.array_equal(ta[3], tb[3]) and ta[4] == tb[4]

    # lifecycle of a restored geometry: transforms, delete, update! all behave like a pushed one
    b.update_transforms(gb, list(W.random_trs(12, seed=9, extent=6.0)))
    b.sync()
    assert b.last_sync_action == rc.RC_SYNC_REFIT
    assert b.delete(gs) and not b.delete(gs)
    b.update(gb, W.box_mesh())
    b.sync()
    assert b.n_geometries() == 1 and b.n_instances() == 12
    a.free()
    b.free()


def test_single_triangle_and_size_query():
    tri = np.array([[0, 0, 1, 1, 0, 1, 0, 1, 1]], F)
    a = rc.TLAS(keep_bvh2=True)
    h = a.push(tri)
    a.sync()
    size = C.c_uint64()
    assert a._lib.rc_export_geometry(a._ctx, h.id, None, 0, C.byref(size)) == L.RC_OK
    blob = a.export_geometry(h)
    assert size.value == blob.nbytes == 128 + 64 + 128 + 64 + 512  # header, 1 BVH2 node, 2 wide slots, 1 triangle (48 -> 64), hull
    small = np.zeros(blob.nbytes - 1, np.uint8)
    assert a._lib.rc_export_geometry(a._ctx, h.id, small.ctypes.data, small.nbytes, C.byref(size)) == L.RC_ERR_INVALID_ARGUMENT
    b = rc.TLAS(keep_bvh2=True)
    b.push_exported(blob)
    b.sync()
    rays = np.zeros(2, RAY_DTYPE)
    rays["o"] = [[0.2, 0.2, 0], [2, 2, 0]]
    rays["d"] = [0, 0, 1]
    rays["t_max"] = np.inf
    ra, rb = a.trace_closest(rays), b.trace_closest(rays)
    assert ra.tobytes() == rb.tobytes() and list(rb["hit"]) == [1, 0] and abs(rb["t"][0] - 1.0) < 1e-6
    with pytest.raises(rc.RaycoreError) as e:
        a.export_geometry(rc.TLASHandle(77))
    assert e.value.code == L.RC_ERR_INVALID_HANDLE
    a.delete(h)
    with pytest.raises(rc.RaycoreError) as e:
        a.export_geometry(h)
    assert e.value.code == L.RC_ERR_DELETED_HANDLE


def test_damaged_blobs_are_refused():
    a = rc.TLAS(keep_bvh2=True)
    h = a.push(np.concatenate([W.box_mesh(), np.zeros((1, 9), F)]))  # 13 submitted faces, 12 kept
    a.sync()
    blob = a.export_geometry(h)
    b = rc.TLAS(keep_bvh2=True)

    def refused(x, what):
        with pytest.raises(rc.RaycoreError) as e:
            b.push_exported(x)
        assert e.value.code == L.RC_ERR_INVALID_ARGUMENT and what in str(e.value), str(e.value)

    refused(blob[:100], "too small")
    refused(blob[:-64], "truncated")
    x = blob.copy(); x[0] = ord("X")
    refused(x, "not a raycore BLAS blob")
    x = blob.copy(); x[8] ^= 0xFF  # layout version
    refused(x, "incompatible")
    x = blob.copy(); x[300] ^= 1  # payload bit flip
    refused(x, "hash mismatch")
    x = blob.copy(); x[:128].view(T.BLOB_HEADER_DTYPE)["n"] += 1  # n without the matching section table
    refused(x, "section table")
    # a self-consistent blob (hash recomputed) whose wide root points outside the node array fails the structural check
    hd = T.blob_header(blob)
    x = blob.copy()
    root = int(hd["off_nodes4"]) + 64
    x[root + 40:root + 44] = np.frombuffer(np.uint32(int(hd["n"]) + 5).tobytes(), np.uint8)  # RcNode4.child0
    x[:128].view(T.BLOB_HEADER_DTYPE)["payload_hash"] = blob_hash(x[128:].tobytes())
    refused(x, "structural check")
    # nothing was appended by the failed imports, and the context still works
    assert b.n_geometries() == 0 and b.n_total_instances() == 0
    b.push_exported(blob)
    b.sync()
    rays = W.box_rays(1000, 3, half=2.0)
    assert b.trace_closest(rays).tobytes() == a.trace_closest(rays).tobytes()
