"""The N>1 path on CPU: world_size-2 ``gloo`` process group, each rank holding its own row chunks, the
kernels replaced by the TEST-ONLY oracle backend.  Checks that the single per-iteration all-reduce of
[k*d sums | k counts | inertia] makes every rank follow exactly the single-process trajectory
(the analogue of the reference's threads/distributed scheduler parametrisation, tests/conftest.py:143-148)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dask_ml_b200 import ChunkedArray
        from dask_ml_b200.cluster import KMeans, k_means as km
        from oracle_backend import OracleBackend

        km._BACKEND_FACTORY = OracleBackend
        rng = np.random.RandomState(7)
        cent = rng.uniform(-10, 10, size=(12, 6))
        X = (cent[rng.randint(0, 12, size=4000)] + rng.standard_normal((4000, 6))).astype(np.float32)
        init = X[:9].copy()
        lo, hi = (0, 1700) if rank == 0 else (1700, 4000)          # uneven shards
        Xl = ChunkedArray.from_array(X[lo:hi], 600)
        a = KMeans(9, init=init, max_iter=12, tol=1e-4).fit(Xl)
        b = KMeans(9, init="k-means||", random_state=3, oversampling_factor=6, max_iter=5).fit(Xl)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), centers=a.cluster_centers_, inertia=a.inertia_,
                 n_iter=a.n_iter_, labels=a.labels_.compute(), centers_b=b.cluster_centers_, inertia_b=b.inertia_)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_ranks_follow_the_single_process_trajectory(tmp_path, oracle, monkeypatch):
    port = _free_port()
    mp.start_processes(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")

    rng = np.random.RandomState(7)
    cent = rng.uniform(-10, 10, size=(12, 6))
    X = (cent[rng.randint(0, 12, size=4000)] + rng.standard_normal((4000, 6))).astype(np.float32)
    init = X[:9].copy()
    lab, inertia, C, n_iter = oracle.kmeans_single_lloyd(oracle.to_blocks(X, 1000), 9, init=init, max_iter=12, tol=1e-4)
    for r in (r0, r1):
        assert int(r["n_iter"]) == n_iter
        np.testing.assert_allclose(r["centers"], C, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(float(r["inertia"]), inertia, rtol=1e-9)
    np.testing.assert_array_equal(np.concatenate([r0["labels"], r1["labels"]]), np.concatenate(lab))
    # both ranks agree bit for bit (same all-reduced buffer -> same centres)
    np.testing.assert_array_equal(r0["centers"], r1["centers"])
    np.testing.assert_array_equal(r0["centers_b"], r1["centers_b"])

    # k-means|| across ranks == k-means|| in one process (draws keyed by the global row index)
    from dask_ml_b200 import ChunkedArray
    from dask_ml_b200.cluster import KMeans, k_means as km
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_backend import OracleBackend

    monkeypatch.setattr(km, "_BACKEND_FACTORY", OracleBackend)
    one = KMeans(9, init="k-means||", random_state=3, oversampling_factor=6, max_iter=5).fit(ChunkedArray.from_array(X, 900))
    np.testing.assert_allclose(r0["centers_b"], one.cluster_centers_, rtol=1e-5, atol=1e-6)
