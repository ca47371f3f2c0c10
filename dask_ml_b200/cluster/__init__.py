"""Unsupervised clustering — the KMeans hot path (dask_ml/cluster/__init__.py:3-5)."""
from .k_means import KMeans, k_means  # noqa: F401
