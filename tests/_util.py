"""Helpers shared by the parity tests."""
import numpy as np


def d2_f64(X, C):
    X = np.asarray(X, dtype=np.float64)
    C = np.asarray(C, dtype=np.float64)
    return ((X[:, None, :] - C[None, :, :]) ** 2).sum(-1) if X.shape[0] * C.shape[0] * X.shape[1] < 5e7 else \
        np.maximum((X * X).sum(1)[:, None] - 2 * X @ C.T + (C * C).sum(1)[None, :], 0)


def assert_labels_match(got, want, X, C, rtol=1e-9, max_frac=1e-3):
    """Labels must be identical except on float64 near-ties: rows where the two chosen centres are
    equidistant to `rtol` relative to (||x||^2+||c||^2).  This is the documented tie-breaking."""
    got = np.asarray(got).astype(np.int64)
    want = np.asarray(want).astype(np.int64)
    assert got.shape == want.shape
    bad = np.nonzero(got != want)[0]
    if len(bad) == 0:
        return 0
    Xb = np.asarray(X, dtype=np.float64)[bad]
    C = np.asarray(C, dtype=np.float64)
    dg = ((Xb - C[got[bad]]) ** 2).sum(1)
    dw = ((Xb - C[want[bad]]) ** 2).sum(1)
    scale = (Xb ** 2).sum(1) + (C ** 2).sum(1).max()
    rel = np.abs(dg - dw) / scale
    assert rel.max() <= rtol, "label mismatch that is not a float64 near-tie: rel margin %g at row %d" % (
        rel.max(), bad[rel.argmax()])
    assert len(bad) <= max(1, max_frac * len(got)), "%d near-tie mismatches of %d rows" % (len(bad), len(got))
    return len(bad)
