"""Parity rules between the CUDA path (or its host simulation) and the oracle — the bar of BASELINE.json:
triangle id, instance id and hit/miss bit-exact except documented near-tie cases (|dt| <= 1e-6 t); t and
barycentrics within 1e-5 relative (in fact bit-identical whenever the same triangle wins, because both sides
evaluate the same Moeller-Trumbore expression in the same instance space).

Classes (DESIGN.md "Parity classes"):
  exact     all fields bit-identical
  tie       both hit, different triangle, |t_a - t_b| <= 1e-6 * max(|t|)          (shared edges/vertices, coplanar duplicates)
  graze     test side found a hit the oracle does not report or a strictly closer one: the reference's slab test is
            not conservative (src/instanced-bvh.jl:1848-1858) and culled a box whose triangle the exact triangle test accepts.
            Verified per ray by re-running the oracle's own triangle test on the reported triangle.
  nan       oracle returned t = NaN (ray lying in a triangle's plane); the wide path rejects NaN hits
  bad       anything else -> test failure
"""
import numpy as np

FIELDS = ("hit", "primitive_id", "instance_id", "instance_custom_index", "meta")


def classify(test_hits, orc_hits, verify_graze=None, tie_rel=1e-6):
    """Returns dict of index arrays per class.  verify_graze(idx) -> bool array (the test side's triangle is a true
    exact-MT hit at exactly the reported t)."""
    a, b = test_hits, orc_hits
    same_ids = np.ones(len(a), bool)
    for f in FIELDS:
        same_ids &= a[f] == b[f]
    bits_equal = same_ids.copy()
    for f in ("t", "bary_u", "bary_v"):
        bits_equal &= a[f].view(np.uint32) == b[f].view(np.uint32)
    exact = bits_equal
    rest = ~exact
    nan = rest & np.isnan(b["t"])
    rest &= ~nan
    both = rest & (a["hit"] == 1) & (b["hit"] == 1)
    with np.errstate(invalid="ignore"):
        tie = both & (np.abs(a["t"] - b["t"]) <= tie_rel * np.maximum(np.abs(a["t"]), np.abs(b["t"])))
    rest &= ~tie
    with np.errstate(invalid="ignore"):
        closer = rest & (a["hit"] == 1) & ((b["hit"] == 0) | (a["t"] < b["t"]))
    graze = np.zeros(len(a), bool)
    if closer.any() and verify_graze is not None:
        idx = np.nonzero(closer)[0]
        ok = verify_graze(idx)
        graze[idx[ok]] = True
    rest &= ~graze
    return {"exact": np.nonzero(exact)[0], "tie": np.nonzero(tie)[0], "graze": np.nonzero(graze)[0], "nan": np.nonzero(nan)[0], "bad": np.nonzero(rest)[0]}


def summarize(cls, n):
    return {k: int(len(v)) for k, v in cls.items()} | {"n": int(n)}


def make_graze_verifier(oracle_module, rays, test_hits, instances, blas_tris_by_index):
    """verify that the test side's reported triangle passes the oracle's exact triangle test at the reported t.
    instances: INSTANCE_DTYPE array (positions = instance_id); blas_tris_by_index: {blas_index(1-based): filtered TRI array in input order}."""

    def verify(idx):
        ok = np.zeros(len(idx), bool)
        for k, i in enumerate(idx):
            h = test_hits[i]
            inst = instances[int(h["instance_id"])]
            tri = blas_tris_by_index[int(inst["blas_index"])][int(h["primitive_id"])]
            d = rays["d"][i].copy()
            d[d == 0] = 0.0
            o = oracle_module.transform_point(inst["inv_transform"], rays["o"][i])
            dd = oracle_module.transform_direction(inst["inv_transform"], d)
            v = tri["v"]
            hit, t, u, vv = oracle_module.intersect_triangle(o, dd, v[0:3], v[3:6], v[6:9], float(rays["t_min"][i]), np.inf)
            ok[k] = hit and np.float32(t) == h["t"] and np.float32(u) == h["bary_u"] and np.float32(vv) == h["bary_v"]
        return ok

    return verify


def assert_parity(cls, n, max_tie_frac=2e-3, max_graze_frac=2e-5, label="", max_graze=None, max_nan=None):
    """max_graze / max_nan: absolute caps.  The BASELINE configurations (C1 / C2 / C3 samples) pass max_graze=0, max_nan=0 — the measured
    value on every committed sample — so that a regression of the conservative-slab slack (more box hits than the reference's slab test
    reports) or of the NaN rule shows up as a failure instead of disappearing in a tolerance."""
    s = summarize(cls, n)
    assert s["bad"] == 0, f"{label}: unexplained mismatches {s}; first: {cls['bad'][:5]}"
    assert s["tie"] <= max(2, max_tie_frac * n), f"{label}: too many ties {s}"
    assert s["graze"] <= (max(1, max_graze_frac * n) if max_graze is None else max_graze), f"{label}: too many graze cases {s}"
    if max_nan is not None:
        assert s["nan"] <= max_nan, f"{label}: NaN-class rays {s}"
    return s
