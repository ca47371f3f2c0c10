"""Host-side helpers on the KMeans path (mirrors dask_ml/utils.py:50-55, 87-158, 295-339)."""
import contextlib
import functools
import logging
from timeit import default_timer as tic

import numpy as np
import sklearn.utils.extmath as skm
import sklearn.utils.validation as sk_validation

from .chunked import ChunkedArray, as_chunked, is_dask_array, is_dask_dataframe, _is_torch

logger = logging.getLogger(__name__)


def row_norms(X, squared=False):
    """Row-wise (squared) euclidean norm; block-wise for chunked input (utils.py:50-55)."""
    if isinstance(X, np.ndarray):
        return skm.row_norms(X, squared=squared)
    X = as_chunked(X)
    from .chunked import block_to_numpy

    return ChunkedArray([skm.row_norms(block_to_numpy(b), squared=squared) for b in X.blocks])


def check_array(array, *args, **kwargs):
    """Validate inputs (utils.py:87-158).

    For chunked arrays a small all-ones sample of the same trailing shape and dtype is passed to
    scikit-learn's ``check_array`` so that shape / dtype errors surface exactly like the
    reference's; ndarrays go straight to scikit-learn.
    """
    accept_dask_array = kwargs.pop("accept_dask_array", True)
    preserve_pandas_dataframe = kwargs.pop("preserve_pandas_dataframe", False)
    accept_dask_dataframe = kwargs.pop("accept_dask_dataframe", False)
    accept_unknown_chunks = kwargs.pop("accept_unknown_chunks", False)
    accept_multiple_blocks = kwargs.pop("accept_multiple_blocks", False)

    if is_dask_dataframe(array):
        if not accept_dask_dataframe:
            raise TypeError("This estimator does not support dask dataframes.")
        return array
    if is_dask_array(array):
        if not accept_dask_array:
            raise TypeError
        if not accept_unknown_chunks and np.isnan(array.shape[0]):
            raise TypeError("Cannot operate on Dask array with unknown chunk sizes.")
        if not accept_multiple_blocks and array.ndim > 1 and len(array.chunks[1]) > 1:
            raise TypeError(
                "Chunking is only allowed on the first axis. "
                "Use 'array.rechunk({1: array.shape[1]})' to "
                "rechunk to a single block along the second axis."
            )
        array = as_chunked(array)
    if isinstance(array, ChunkedArray):
        if not accept_dask_array:
            raise TypeError
        shape = array.shape
        if len(shape) == 2:
            shape = (min(10, shape[0]), shape[1])
        else:
            shape = (min(10, shape[0]),)
        sample = np.ones(shape=shape, dtype=array.dtype)
        sk_validation.check_array(sample, *args, **kwargs)
        return array
    if _is_torch(array):
        if array.ndim != 2:
            raise ValueError("Expected 2D array, got %dD tensor instead" % array.ndim)
        sample = np.ones((min(10, array.shape[0]), array.shape[1]))
        sk_validation.check_array(sample, *args, **kwargs)
        return array
    try:
        import pandas as pd

        if isinstance(array, pd.DataFrame) and preserve_pandas_dataframe:
            return array
    except ImportError:  # pragma: no cover
        pass
    return sk_validation.check_array(array, *args, **kwargs)


def _format_bytes(n):
    for unit, div in (("GB", 1e9), ("MB", 1e6), ("kB", 1e3)):
        if n > div:
            return "%0.2f %s" % (n / div, unit)
    return "%d B" % n


def _log_array(logger, arr, name):
    logger.info("%s: %s, %s blocks", name, _format_bytes(arr.nbytes), getattr(arr, "numblocks", "No"))


@contextlib.contextmanager
def _timer(name, _logger=None, level="info"):
    """Log the wall time of a block (utils.py:295-326); same message format as the reference."""
    start = tic()
    _logger = _logger or logger
    _logger.info("Starting %s", name)
    nvtx = _nvtx()
    if nvtx is not None:
        nvtx.range_push(str(name))         # the phases of fit show up as ranges in an Nsight Systems / ncu --nvtx timeline
    try:
        yield
    finally:
        if nvtx is not None:
            nvtx.range_pop()
    stop = tic()
    delta = stop - start
    getattr(_logger, level)("Finished %s in %0.4fs", name, delta)


def _nvtx():
    """torch.cuda.nvtx when a CUDA context exists in this process (never initialises one), else None."""
    try:
        import torch
        if torch.cuda.is_available() and torch.cuda.is_initialized():
            return torch.cuda.nvtx
    except Exception:
        pass
    return None


def _timed(_logger=None, level="info"):
    """Decorator form of ``_timer`` (utils.py:329-339)."""

    def wrapper(func):
        @functools.wraps(func)
        def inner(*args, **kwargs):
            with _timer(func.__name__, _logger=_logger, level=level):
                return func(*args, **kwargs)

        return inner

    return wrapper
