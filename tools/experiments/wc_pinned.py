#!/usr/bin/env python
"""Experiment: does write-combined page-locked memory (cudaHostAllocWriteCombined) speed up the H2D leg of the host-buffer pipeline on this
box?  Times 1 GiB H2D copies from default and from write-combined pinned memory, alone and with a concurrent 1 GiB D2H."""
import ctypes as C
import json
import torch

rt = C.CDLL("libcudart.so.12")
torch.cuda.init()
torch.zeros(1, device="cuda")
N = 1 << 30


def host_alloc(flags):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(N), C.c_uint(flags)) == 0
    C.memset(p, 1, N)
    return p


def ev():
    e = C.c_void_p()
    assert rt.cudaEventCreate(C.byref(e)) == 0
    return e


s1, s2 = C.c_void_p(), C.c_void_p()
rt.cudaStreamCreate(C.byref(s1)); rt.cudaStreamCreate(C.byref(s2))
d_in = torch.empty(N, dtype=torch.uint8, device="cuda")
d_out = torch.empty(N, dtype=torch.uint8, device="cuda")
h_out = host_alloc(0)
out = {}
for name, flags in (("default", 0), ("write_combined", 4)):
    h = host_alloc(flags)
    for both in (False, True):
        best = 1e9
        for _ in range(4):
            e0, e1 = ev(), ev()
            rt.cudaDeviceSynchronize()
            rt.cudaEventRecord(e0, s1)
            rt.cudaMemcpyAsync(C.c_void_p(d_in.data_ptr()), h, C.c_size_t(N), 1, s1)
            if both:
                rt.cudaMemcpyAsync(h_out, C.c_void_p(d_out.data_ptr()), C.c_size_t(N), 2, s2)
            rt.cudaEventRecord(e1, s1)
            rt.cudaDeviceSynchronize()
            ms = C.c_float()
            rt.cudaEventElapsedTime(C.byref(ms), e0, e1)
            best = min(best, ms.value)
        out[f"{name}_h2d_GBs" + ("_with_d2h" if both else "")] = N / best / 1e6
print(json.dumps(out))
