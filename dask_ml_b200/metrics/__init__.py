"""Distance operators on the KMeans path (dask_ml/metrics/pairwise.py:18-97)."""
from .pairwise import (  # noqa: F401
    euclidean_distances,
    pairwise_distances,
    pairwise_distances_argmin_min,
)
