"""Distance operators on the KMeans path (dask_ml/metrics/pairwise.py:18-97) and the kernel built on them (:131-139)."""
from .pairwise import (  # noqa: F401
    euclidean_distances,
    pairwise_distances,
    pairwise_distances_argmin_min,
    pairwise_kernels,
    rbf_kernel,
)
