import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch, ctypes as C
import raycore_b200 as rc
from raycore_b200 import workloads as W
n = 1 << 24
tl = rc.TLAS(); tl.push(W.bumpy_sphere(709)); tl.sync()
rays = W.interior_rays(n, 77, radius=0.8)
h_r = torch.from_numpy(rays.view(np.uint8).reshape(-1)).pin_memory(); h_h = torch.empty(n*32, dtype=torch.uint8).pin_memory()
d = torch.empty(n*32, dtype=torch.uint8, device='cuda'); d2 = torch.empty(n*32, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps
def h2d():
    with torch.cuda.stream(s1): d.copy_(h_r, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_h.copy_(d2, non_blocking=True)
def both(): h2d(); d2h()
gb = n*32/1e9
print('H2D GB/s', gb/t(h2d), 'D2H GB/s', gb/t(d2h), 'both (each) GB/s', gb/t(both))
lib, ctx = tl._lib, tl._ctx
def e2e(): assert lib.rc_trace_closest(ctx, h_r.data_ptr(), h_h.data_ptr(), n, 0) == 0
print('chunk', os.environ.get('RC_HOST_CHUNK_RAYS'), 'e2e Mrays/s', n/t(e2e)/1e6)
