"""Profiling helper: one max-pool launch shape repeatedly.  python tests/tools/run_one_pool.py C T H W k s p [B]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from coclr_b200 import ops, lib as L  # noqa: E402

Cc, T, H, W, k, s, p = [int(v) for v in sys.argv[1:8]]
B = int(sys.argv[8]) if len(sys.argv) > 8 else 32
kk = (k, k, k) if k != 1 else (1, 3, 3)
ss = (s, s, s) if k != 1 else (1, s, s)
pp = (p, p, p) if k != 1 else (0, p, p)
geom = ops.Geometry(kk, ss, pp)
To, Ho, Wo = geom.out_dims(T, H, W)
x = ops.Planes((B, T, H, W, Cc), 0, "cuda"); x.hi.normal_().abs_(); x.lo.normal_(0, 1e-3)
y = ops.Planes((B, To, Ho, Wo, Cc), 0, "cuda")
tw = ops.Planes((B, To, Ho, Wo, Cc), 1, "cuda")
idx = torch.empty(B * To * Ho * Wo * Cc, dtype=torch.uint8, device="cuda")
dy = torch.randn(B, To, Ho, Wo, Cc, device="cuda")
dx = torch.zeros(B, T, H, W, Cc, device="cuda")
pl = L.Pool(L.dptr(x.hi), L.dptr(x.lo), Cc, 0, L.dptr(y.hi), L.dptr(y.lo), Cc, 0, L.dptr(tw.hi), L.dptr(tw.lo), L.dptr(idx),
            B, Cc, T, H, W, To, Ho, Wo, geom.c(0), L.dptr(dy), L.dptr(dx), 0)
lib = L.load()
for name, fn in (("fwd", lib.coclr_maxpool_fwd), ("bwd", lib.coclr_maxpool_bwd)):
    for _ in range(2):
        fn(C.byref(pl), L.stream_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn(C.byref(pl), L.stream_ptr())
    e1.record(); torch.cuda.synchronize()
    print("pool %s %s: %.3f ms" % (name, sys.argv[1:8], e0.elapsed_time(e1) / 5))
