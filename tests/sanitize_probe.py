"""Manual probe (not collected by pytest): one small call of every kernel family, for compute-sanitizer.
    compute-sanitizer --tool memcheck|racecheck|synccheck python tests/sanitize_probe.py [family ...]
Families: tc (all four tc_chunk_kernel variants + transform), tc2, stream, simt, aux."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dask_ml_b200.engine import CudaBackend

fams = sys.argv[1:] or ["tc", "tc2", "stream", "simt", "aux"]
be = CudaBackend()
g = torch.Generator(device="cuda").manual_seed(0)


def data(n, d, k, dtype=torch.float32):
    cent = torch.empty((max(2, k // 2), d), device="cuda").uniform_(-10, 10, generator=g)
    X = (cent[torch.randint(0, cent.shape[0], (n,), device="cuda", generator=g)] + torch.randn((n, d), device="cuda", generator=g)).to(dtype)
    x = be.to_device(X, dtype)
    C = X[:k].double().contiguous()
    return x, C


def run(x, C, dtype, flags=0, want_dist=True):
    n, d = x.shape
    k = C.shape[0]
    be.flags = flags
    pack = be.pack_centers(C, dtype)
    out_dt = torch.float32 if dtype == torch.bfloat16 else dtype
    lab = be.empty((n,), torch.int32); md = be.empty((n,), out_dt)
    sums = be.zeros((k * d,), torch.float64); cnt = be.zeros((k,), torch.int64); acc = be.zeros((1,), torch.float64)
    be.lloyd_chunk(x, pack, k, lab, None, sums, cnt, None)              # pure Lloyd variant
    be.lloyd_chunk(x, pack, k, lab, md, sums, cnt, acc)                 # sums + distances
    be.assign_chunk(x, pack, k, lab, None, True, None)                  # labels only
    be.assign_chunk(x, pack, k, lab, md, False, acc)                    # labels + distances
    torch.cuda.synchronize()
    be.flags = 0
    return int(cnt.sum())


if "tc" in fams:
    for (n, d, k) in ((1500, 64, 256), (900, 41, 100)):
        x, C = data(n, d, k)
        print("tc", n, d, k, run(x, C, torch.float32, flags=2))
        out = be.empty((n, k), torch.float32)
        be.transform_chunk(x, be.pack_centers(C, torch.float32), k, out)
        torch.cuda.synchronize()
if "tc2" in fams:
    for (n, d, k) in ((1500, 128, 1024), (700, 64, 300)):
        x, C = data(n, d, k, torch.bfloat16)
        print("tc2", n, d, k, run(x, C, torch.bfloat16))
if "stream" in fams:
    for (n, d, k) in ((5000, 13, 20), (3000, 16, 31)):
        x, C = data(n, d, k)
        print("stream", n, d, k, run(x.contiguous() if d % 4 else x, C, torch.float32))
if "simt" in fams:
    for (n, d, k, dt) in ((2000, 24, 40, torch.float32), (1500, 16, 8, torch.float64), (300, 300, 50, torch.float32)):
        x, C = data(n, d, k, dt)
        print("simt", n, d, k, run(x, C, dt, flags=1))
if "aux" in fams:
    from dask_ml_b200.cluster import KMeans
    X = torch.randn((20000, 8), device="cuda", generator=g)
    km = KMeans(5, init="k-means||", random_state=0, oversampling_factor=10, max_iter=10).fit(X)
    print("aux fit", km.n_iter_, float(km.inertia_))
print("done")
