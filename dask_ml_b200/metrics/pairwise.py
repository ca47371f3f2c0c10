"""GPU versions of the three distance operators the KMeans path uses
(dask_ml/metrics/pairwise.py:18-52, 55-66, 69-97).  Only the euclidean metric exists on this
path; every result is a device-resident ``ChunkedArray`` with the reference's dtypes."""
import numpy as np
import torch

from ..chunked import ChunkedArray


def _prep(X, Y):
    from ..cluster.k_means import _to_device_data

    X = _to_device_data(X, check_finite=False)
    Y = np.asarray(Y)
    if Y.ndim != 2 or Y.shape[1] != X.d:
        raise ValueError(
            "Incompatible dimension for X and Y matrices: X.shape[1] == %d while Y.shape[1] == %d"
            % (X.d, Y.shape[1] if Y.ndim == 2 else -1)
        )
    be = X.backend
    C = torch.as_tensor(np.ascontiguousarray(Y, dtype=np.float64)).to(be.device)
    pack = be.pack_centers(C, X.dtype)
    return X, be, pack, int(Y.shape[0])


def pairwise_distances_argmin_min(X, Y, axis=1, metric="euclidean", batch_size=None, metric_kwargs=None):
    """Per row of X: index of and distance to the nearest row of Y (pairwise.py:18-52).

    Returns ``(argmins int64, mins float64)`` like the reference (pairwise.py:41-49).  ``mins`` is
    the euclidean distance, or its square with ``metric_kwargs={'squared': True}``.
    """
    if type(Y).__module__.startswith("dask"):
        raise TypeError("`Y` must be a numpy array")
    if metric not in ("euclidean", "sqeuclidean", "l2"):
        raise NotImplementedError("only the euclidean metric is on the B200 KMeans path, got %r" % (metric,))
    if axis != 1:
        raise NotImplementedError("axis must be 1")
    squared = bool((metric_kwargs or {}).get("squared", False)) or metric == "sqeuclidean"
    X, be, pack, k = _prep(X, Y)
    acc = be.zeros((1,), torch.float64)
    argmins, mins = [], []
    for x in X.chunks:
        n = int(x.shape[0])
        lab = be.empty((n,), torch.int32)
        mn = be.empty((n,), X.out_dtype)
        be.assign_chunk(x, pack, k, lab, mn, squared, acc)
        argmins.append(lab.to(torch.int64))
        mins.append(mn.to(torch.float64))
    return ChunkedArray(argmins), ChunkedArray(mins)


def pairwise_distances(X, Y, metric="euclidean", n_jobs=None, **kwargs):
    """Full (n, len(Y)) euclidean distance blocks (pairwise.py:55-66)."""
    if type(Y).__module__.startswith("dask"):
        raise TypeError("`Y` must be a numpy array")
    if metric != "euclidean":
        raise NotImplementedError("only the euclidean metric is on the B200 KMeans path, got %r" % (metric,))
    return euclidean_distances(X, Y)


def euclidean_distances(X, Y=None, Y_norm_squared=None, squared=False, X_norm_squared=None):
    """sqrt(max(||x||^2 - 2 x.y + ||y||^2, 0)) for every (row of X, row of Y)  (pairwise.py:69-97).

    The result has the dtype of X.  ``X_norm_squared`` / ``Y_norm_squared`` are validated for
    shape as in the reference but the kernel recomputes the norms on the fly (they are free).
    """
    if Y is None:
        raise NotImplementedError("Y=None (X against itself) is not on the KMeans path")
    from ..cluster.k_means import _to_device_data

    X = _to_device_data(X, check_finite=False)
    Y = np.asarray(Y)
    if X.dtype == torch.bfloat16:
        # the full distance matrix of bf16 rows is produced in float32 (no bf16 transform kernel)
        from ..engine import DeviceData

        X = DeviceData([c.to(torch.float32) for c in X.chunks], X.backend, X.comm)
    if X.dtype == torch.float32 and Y.dtype == np.float64:
        # result dtype follows numpy promotion of (X, Y) like -2*dot(X, Y.T)+XX+YY in the reference
        from ..engine import DeviceData

        X = DeviceData([c.to(torch.float64) for c in X.chunks], X.backend, X.comm)
    X, be, pack, k = _prep(X, Y)
    if X_norm_squared is not None:
        XX = np.asarray(X_norm_squared)
        if XX.shape not in ((1, X.n_local), (X.n_local, 1)):
            raise ValueError("Incompatible dimensions for X and X_norm_squared")
    if Y_norm_squared is not None:
        YY = np.asarray(Y_norm_squared)
        if YY.ndim < 2:
            YY = YY[:, np.newaxis]
        if YY.shape != (1, k):
            raise ValueError("Incompatible dimensions for Y and Y_norm_squared")
    outs = []
    for x in X.chunks:
        out = be.empty((int(x.shape[0]), k), X.dtype)
        be.transform_chunk(x, pack, k, out)
        outs.append(out * out if squared else out)
    return ChunkedArray(outs)
