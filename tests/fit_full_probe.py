import time, sys, os, logging
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from bench import synth_blobs_device
from dask_ml_b200.cluster import KMeans
X = synth_blobs_device(10_000_000, 64, 256, 5, torch.device("cuda:0"), torch.float32)
torch.cuda.synchronize()
for init, kw in (("k-means||", dict(oversampling_factor=2)), ("k-means||", dict(oversampling_factor=64))):
    t0 = time.perf_counter()
    km = KMeans(n_clusters=256, init=init, random_state=0, max_iter=50, **kw).fit(X)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    p = km.predict(X); torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("fit 10M x 64 k=256 init=%s %s: %.2f s, n_iter=%d, inertia_=%.6g ; predict %.3f s" % (init, kw, t1 - t0, km.n_iter_, km.inertia_, t2 - t1), flush=True)
