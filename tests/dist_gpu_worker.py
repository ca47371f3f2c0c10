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
        # the peer-memory all-reduce (bkm_p2p.cu) against the sum in rank order, many calls in a row (slot parity reuse)
        from dask_ml_b200.engine import Comm, _p2p_state
        comm = Comm()
        st = _p2p_state(comm)
        p2p_ready = st is not None
        p2p_bad = 0
        g = torch.Generator(device="cuda").manual_seed(1234 + rank)
        for i, m in enumerate([1, 7, 16641, 1000, 65536, 3, 16641, 16641, 40000]):
            t = torch.randn(m, dtype=torch.float64, device="cuda", generator=g) * (10.0 ** (i % 5))
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t)
            want = torch.zeros_like(t)
            for q in range(world):
                want += parts[q]                      # rank order: what the kernel does
            got = comm.allreduce_sum_(t.clone())
            p2p_bad += int((got != want).sum())
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), centers=a.cluster_centers_, inertia=a.inertia_,
                 n_iter=a.n_iter_, labels=a.labels_.compute(), lo=bounds[rank], hi=bounds[rank + 1],
                 centers_b=b.cluster_centers_, p2p_ready=p2p_ready, p2p_bad=p2p_bad)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
