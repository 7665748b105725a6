"""Serialised-geometry blobs without a GPU: the library's host-side import checks (rc_check_exported) against blobs assembled here from
the host simulation of the builder — header layout shared between C (RcBlobHeader) and Python (BLOB_HEADER_DTYPE), section table,
payload hash, refusal of damaged input.  The GPU round trip itself is tests/test_gpu_export.py."""
import numpy as np
import pytest

import raycore_b200 as rc
from raycore_b200 import _lib as L
from raycore_b200 import tlas as T
from raycore_b200 import workloads as W
import hostsim_py as hs


def _up64(x):
    return (x + 63) & ~63


def make_blob(verts, normals=False, bvh2=True):
    """A blob with the library's layout: header [+ BVH2 nodes (64 B), only for geometry built with RC_BUILD_KEEP_BVH2] + wide-node slots
    (zero here) + sorted triangles + hull (+ normals)."""
    b = hs.HsBlas(verts)
    n = b.n
    order = b.order()
    v = np.asarray(verts, np.float32).reshape(-1, 9)
    keep = np.nonzero(~W.is_degenerate(v))[0]
    hdr = np.zeros(1, T.BLOB_HEADER_DTYPE)
    o = 128
    if bvh2:
        hdr["off_nodes2"] = o; o = _up64(o + 64 * (2 * n - 1))
    hdr["off_nodes4"] = o; o = _up64(o + 64 * (n + 1))
    hdr["off_tris"] = o; o = _up64(o + 48 * n)
    hdr["off_hull"] = o; o = _up64(o + 32 * 16)
    if normals:
        hdr["off_normals"] = o; o = _up64(o + 36 * n)
    hdr["magic"], hdr["abi_version"], hdr["leaf_max"], hdr["hull_boxes"] = b"RCBLAS\x00\x02", 1, 2, 16
    hdr["n"], hdr["n_faces_in"], hdr["has_normals"], hdr["total_bytes"] = n, len(v), int(normals), o
    hdr["root_aabb"] = b.root()
    blob = np.zeros(o, np.uint8)
    if bvh2:
        nodes = np.zeros((2 * n - 1, 64), np.uint8)
        nodes[:, :60] = b.nodes2().view(np.uint8).reshape(-1, 60)
        blob[int(hdr["off_nodes2"][0]):][: nodes.size] = nodes.reshape(-1)
    tri = np.zeros(n, T.BLOB_TRI_DTYPE)
    src = v[keep][order]
    tri["v0"], tri["v1"], tri["v2"] = src[:, 0:3], src[:, 3:6], src[:, 6:9]
    tri["prim_id"], tri["face_index"], tri["metadata"] = order, keep[order], keep[order] + 1
    blob[int(hdr["off_tris"][0]):][: tri.nbytes] = tri.view(np.uint8)
    hdr["payload_hash"] = T.blob_hash(blob[128:].tobytes())
    blob[:128] = hdr.view(np.uint8)
    return blob


def _refused(blob, what):
    with pytest.raises(rc.RaycoreError) as e:
        T.check_exported(blob)
    assert e.value.code == L.RC_ERR_INVALID_ARGUMENT and what in str(e.value), str(e.value)


@pytest.mark.parametrize("normals", [False, True])
def test_wellformed_blob_is_accepted_and_parsed(normals):
    verts = np.concatenate([W.box_mesh(), np.zeros((1, 9), np.float32), W.uv_sphere(7)])  # one degenerate face in the soup
    blob = make_blob(verts, normals)
    n, f, hn = T.check_exported(blob)
    keep = np.nonzero(~W.is_degenerate(verts))[0]  # the sphere's pole faces are degenerate too
    assert (n, f, hn) == (len(keep), len(verts), normals) and 12 not in keep
    assert T.check_exported(blob.tobytes()) == (n, f, hn)  # bytes and arrays alike
    h = T.blob_header(blob)
    assert h["n"] == n and h["total_bytes"] == blob.nbytes and (h["off_normals"] > 0) == normals
    faces = T.blob_faces(blob)  # the submitted soup, face for face (the dropped face stays zero)
    dropped = np.setdiff1d(np.arange(len(verts)), keep)
    assert np.array_equal(faces[keep], verts[keep]) and not faces[dropped].any()
    assert sorted(T.blob_triangles(blob)["prim_id"].tolist()) == list(range(n))


def test_single_triangle_blob_size():
    tri = np.array([[0, 0, 1, 1, 0, 1, 0, 1, 1]], np.float32)
    blob = make_blob(tri)
    assert blob.nbytes == 128 + 64 + 128 + 64 + 512 and T.check_exported(blob) == (1, 1, False)
    # the default build carries no reference-layout BVH2: the section is simply absent (off_nodes2 == 0)
    blob = make_blob(tri, bvh2=False)
    assert blob.nbytes == 128 + 128 + 64 + 512 and T.check_exported(blob) == (1, 1, False) and T.blob_header(blob)["off_nodes2"] == 0


def test_damaged_blobs_are_refused_on_the_host():
    blob = make_blob(np.concatenate([W.box_mesh(), np.zeros((1, 9), np.float32)]))
    _refused(blob[:0], "too small")
    _refused(blob[:100], "too small")
    _refused(blob[:-64], "truncated")
    x = blob.copy(); x[0] = ord("X")
    _refused(x, "not a raycore BLAS blob")
    x = blob.copy(); x[:128].view(T.BLOB_HEADER_DTYPE)["abi_version"] += 1
    _refused(x, "incompatible")
    x = blob.copy(); x[:128].view(T.BLOB_HEADER_DTYPE)["leaf_max"] = 4
    _refused(x, "incompatible")
    x = blob.copy(); x[300] ^= 1
    _refused(x, "hash mismatch")
    x = blob.copy(); x[-1] ^= 0x80  # the last payload byte is hashed too
    _refused(x, "hash mismatch")
    x = blob.copy(); x[:128].view(T.BLOB_HEADER_DTYPE)["n"] += 1
    _refused(x, "section table")
    x = blob.copy(); x[:128].view(T.BLOB_HEADER_DTYPE)["n"] = 0
    _refused(x, "bad triangle count")
    x = blob.copy(); x[:128].view(T.BLOB_HEADER_DTYPE)["n_faces_in"] = 3
    _refused(x, "bad triangle count")
    x = blob.copy(); x[:128].view(T.BLOB_HEADER_DTYPE)["root_aabb"][0, 3] = 3e38; x[:128].view(T.BLOB_HEADER_DTYPE)["root_aabb"][0, 0] = -3e38
    _refused(x, "supported range")
    # trailing bytes after the blob are tolerated (a blob inside a larger file mapping)
    assert T.check_exported(np.concatenate([blob, np.zeros(77, np.uint8)]))[0] == 12
