"""The sort of the one-kernel builders (k_build_small / k_tlas_small, raycore.jl_b200/csrc/rc_build.cu `sb_sort_runs`), restated in numpy:
runs of 256 keys are sorted by counting (rank = keys of the run that are smaller, or equal and earlier), then every key is placed at
rank-in-run + for every other run the number of its keys that are smaller (later runs) or not larger (earlier runs).  The result must be the
stable sort the reference's `AK.sortperm` gives (src/instanced-bvh.jl:1399) — the GPU tests check the kernels byte for byte, this checks
the rule itself, including long runs of equal keys across run boundaries."""
import numpy as np
import pytest

RUN = 256


def run_merge_sort(keys):
    keys = np.asarray(keys, np.uint32)
    n = len(keys)
    runs = [np.arange(a, min(n, a + RUN)) for a in range(0, n, RUN)]
    # S1: counting sort of every run (stable: ties by position)
    sorted_runs = []
    for idx in runs:
        k = keys[idx]
        rank = np.array([np.sum((k < k[i]) | ((k == k[i]) & (np.arange(len(k)) < i))) for i in range(len(k))])
        out_k, out_i = np.empty_like(k), np.empty_like(idx)
        out_k[rank], out_i[rank] = k, idx
        sorted_runs.append((out_k, out_i))
    # S2: merge position = rank in the own run + lower bound in later runs + upper bound in earlier runs
    perm = np.empty(n, np.int64)
    skeys = np.empty(n, np.uint32)
    for r, (k, idx) in enumerate(sorted_runs):
        pos = np.arange(len(k))
        for q, (ko, _) in enumerate(sorted_runs):
            if q == r:
                continue
            pos = pos + np.searchsorted(ko, k, side="right" if q < r else "left")
        perm[pos], skeys[pos] = idx, k
    return skeys, perm


@pytest.mark.parametrize("n", [1, 2, 255, 256, 257, 1000, 4099])
def test_run_merge_sort_is_the_stable_sort(n):
    rs = np.random.RandomState(n)
    for keys in (rs.randint(0, 1 << 30, n), rs.randint(0, 7, n), np.zeros(n, np.int64), np.arange(n)[::-1] // 3):
        keys = keys.astype(np.uint32)
        skeys, perm = run_merge_sort(keys)
        want = np.argsort(keys, kind="stable")
        assert np.array_equal(perm, want)
        assert np.array_equal(skeys, keys[want])
