"""``make_blobs`` with the contract of dask_ml.datasets.make_blobs (dask_ml/datasets.py:76-202).

The reference builds prototype centres from ONE scikit-learn call with the user's seed (the per-cluster means of a
first-block-sized sample, datasets.py:160-176) and then generates every block independently with
``sklearn.datasets.make_blobs(n_block, centers=prototype, random_state=block_index)`` (datasets.py:178-189).

* ``device=None`` (default): exactly that, block by block on the host -> bit-identical to the reference for the
  same scikit-learn version (BASELINE config C1: 100k x 16 float64, 8 blocks).  Returns host ``ChunkedArray``s.
* ``device='cuda'``: the same prototype centres, but every block is generated ON THE GPU by ``bkm_make_blobs_chunk``
  (Philox streams keyed by the block index, so a block is reproducible whatever GPU generates it).  numpy's
  Mersenne-Twister stream cannot be reproduced on the device: the blocks are statistically equivalent, not identical
  (labels are i.i.d. uniform instead of an exact equal split).  Returns device-resident ``ChunkedArray``s.
"""
from numbers import Integral

import numpy as np

from .chunked import ChunkedArray


def _normalize_chunks(chunks, n_samples, n_features):
    """Row block sizes from the forms datasets.py:133-139 accepts (blocksize, blockshape, explicit sizes)."""
    if chunks is None:
        return [int(n_samples)]
    if isinstance(chunks, Integral):
        c = int(chunks)
    elif isinstance(chunks, (tuple, list)) and len(chunks) and isinstance(chunks[0], (tuple, list)):
        if len(chunks) > 1 and len(chunks[1]) > 1:
            raise ValueError("Can only generate arrays partitioned along the first axis. Specifying a larger chunksize "
                             "for the second axis.")
        sizes = [int(v) for v in chunks[0]]
        if sum(sizes) != n_samples:
            raise ValueError("chunks do not add up to n_samples")
        return sizes
    else:
        if len(chunks) > 1 and int(chunks[1]) < n_features:
            raise ValueError("Can only generate arrays partitioned along the first axis. Specifying a larger chunksize "
                             "for the second axis.")
        c = int(chunks[0])
    c = max(1, c)
    return [c] * (n_samples // c) + ([n_samples % c] if n_samples % c else [])


def make_blobs(n_samples=100, n_features=2, centers=None, cluster_std=1.0, center_box=(-10.0, 10.0), shuffle=True,
               random_state=None, chunks=None, device=None, dtype=None):
    """Generate isotropic Gaussian blobs for clustering, one row block at a time (datasets.py:76-202).

    Returns ``(X, y)``: ``ChunkedArray`` of shape (n_samples, n_features) float64 and (n_samples,) int64 like the
    reference (``dtype`` lets the device generator write float32 directly)."""
    import sklearn.datasets

    sizes = _normalize_chunks(chunks, int(n_samples), int(n_features))
    if centers is None:
        centers = 3
    if isinstance(centers, Integral):
        # prototype centres: per-cluster means of one first-block-sized sample drawn with the user's seed
        n_centers = int(centers)
        Xp, yp = sklearn.datasets.make_blobs(n_samples=sizes[0], n_features=n_features, centers=n_centers,
                                             shuffle=shuffle, cluster_std=cluster_std, center_box=center_box,
                                             random_state=random_state)
        centers = np.zeros((n_centers, n_features))
        for i in range(n_centers):
            centers[i] = Xp[yp == i].mean(0)
    centers = np.asarray(centers, dtype=np.float64)
    if device is None:
        Xs, ys = [], []
        for i, m in enumerate(sizes):
            Xb, yb = sklearn.datasets.make_blobs(n_samples=m, n_features=n_features, centers=centers,
                                                 cluster_std=cluster_std, shuffle=shuffle, center_box=center_box,
                                                 random_state=i)
            Xs.append(Xb if dtype is None else Xb.astype(dtype))
            ys.append(yb.astype(np.int64))
        return ChunkedArray(Xs), ChunkedArray(ys)

    import ctypes

    import torch

    from . import _lib
    from .engine import CudaBackend, _DT_CODE

    be = CudaBackend(torch.device(device) if not isinstance(device, torch.device) else device)
    tdt = torch.float64 if dtype is None else {np.dtype("float32"): torch.float32, np.dtype("float64"): torch.float64}[np.dtype(dtype)]
    k = int(centers.shape[0])
    std = np.broadcast_to(np.asarray(cluster_std, dtype=np.float64), (k,)).copy()
    Cd = torch.as_tensor(centers).to(be.device)
    Sd = torch.as_tensor(std).to(be.device)
    Xs, ys = [], []
    with torch.cuda.device(be.device):
        for i, m in enumerate(sizes):
            Xb = torch.empty((m, int(n_features)), dtype=tdt, device=be.device)
            yb = torch.empty((m,), dtype=torch.int64, device=be.device)
            _lib.check(be.lib.bkm_make_blobs_chunk(
                ctypes.c_void_p(Xb.data_ptr()), ctypes.c_void_p(yb.data_ptr()), m, int(n_features), int(n_features),
                _DT_CODE[tdt], ctypes.c_void_p(Cd.data_ptr()), ctypes.c_void_p(Sd.data_ptr()), k, i,
                ctypes.c_void_p(torch.cuda.current_stream(be.device).cuda_stream)), "bkm_make_blobs_chunk")
            Xs.append(Xb)
            ys.append(yb)
    return ChunkedArray(Xs), ChunkedArray(ys)
