"""Manual pipeline-timeline probe (not collected by pytest).  Needs a `make -C dask_ml_b200/csrc TRACE=1`
build: prints, for CTA 0, when each pipeline role reached each event of tiles [lo, hi) in SM cycles.
Run on the GPU box:  python tests/tc_trace.py [lo hi] [dist|lloyd] [d k]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dask_ml_b200.engine import CudaBackend
from dask_ml_b200 import _lib

lo = int(sys.argv[1]) if len(sys.argv) > 1 else 40
hi = int(sys.argv[2]) if len(sys.argv) > 2 else 52
want_dist = len(sys.argv) > 3 and sys.argv[3] == "dist"
be = CudaBackend(flags=_lib.FLAG_FORCE_TC)
d = int(sys.argv[4]) if len(sys.argv) > 4 else 64
k = int(sys.argv[5]) if len(sys.argv) > 5 else 256
n = 148 * 128 * 128
g = torch.Generator(device="cuda").manual_seed(0)
cent = torch.empty((k, d), device="cuda").uniform_(-10, 10, generator=g)
X = be.to_device(cent[torch.randint(0, k, (n,), device="cuda", generator=g)] + torch.randn((n, d), device="cuda", generator=g), torch.float32)
C = X[torch.randperm(n, device="cuda", generator=g)[:k]].double()
pack = be.pack_centers(C, torch.float32)
labels = be.empty((n,), torch.int32)
mind2 = be.empty((n,), torch.float32) if want_dist else None
sums = be.zeros((k * d,), torch.float64); counts = be.zeros((k,), torch.int64); inertia = be.zeros((1,), torch.float64)
for _ in range(2):
    be.lloyd_chunk(X, pack, k, labels, mind2, sums, counts, inertia if want_dist else None)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (16 * 128))()
got = be.lib.bkm_debug_trace(ctypes.cast(buf, ctypes.c_void_p), 16 * 128)
if got == 0:
    sys.exit("library was not built with TRACE=1")
t = np.array(buf[:], dtype=np.int64).reshape(16, 128)
names = ["tma_issue", "conv_start", "conv_end", "mma0_start", "mma0_commit", "mma1_start", "mma1_commit",
         "epi0_start", "epi0_end", "epi1_start", "epi1_end", "labels", "mstep_start", "mstep_end"]
t0 = t[1, lo]
print("tile " + " ".join("%12s" % s for s in names))
for it in range(lo, hi):
    print("%4d " % it + " ".join("%12d" % (t[s, it] - t0) for s in range(len(names))))
per = (t[11, hi - 1] - t[11, lo]) / (hi - 1 - lo)
print("cycles per tile (labels event): %.0f" % per)
print("M-step warp 0: list-walk iterations per tile:", " ".join(str(int(t[14, it])) for it in range(lo, hi)))
