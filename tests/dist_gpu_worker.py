"""Worker for tests/test_gpu_multi.py (launched by torch.distributed.run, one rank per GPU, NCCL)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def main(out_dir):
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        from dask_ml_b200 import ChunkedArray
        from dask_ml_b200.cluster import KMeans

        rng = np.random.RandomState(11)
        n, d, k = 64000, 64, 256
        cent = rng.uniform(-10, 10, size=(90, d))
        X = (cent[rng.randint(0, 90, size=n)] + rng.standard_normal((n, d))).astype(np.float32)
        init = X[:k].copy()
        bounds = np.linspace(0, n, world + 1).astype(int)
        Xl = ChunkedArray.from_array(X[bounds[rank]:bounds[rank + 1]], 9000)
        a = KMeans(k, init=init, max_iter=6, tol=1e-4).fit(Xl)
        b = KMeans(16, init="k-means||", random_state=1, oversampling_factor=20, max_iter=4).fit(Xl)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), centers=a.cluster_centers_, inertia=a.inertia_,
                 n_iter=a.n_iter_, labels=a.labels_.compute(), lo=bounds[rank], hi=bounds[rank + 1],
                 centers_b=b.cluster_centers_)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
