"""The near-tie bound tau (bkm_api.cu: tau_for) is what makes the fp32 / split-fp16 / bf16-split E-step safe: every row
whose two best candidates are closer than tau * (||x||^2 + max||c||^2) is re-decided in float64.  These tests measure,
per kernel family, how wrong the fast arithmetic can be when that re-check is switched OFF (BKM_FLAG_NO_RECHECK): the
largest float64 margin of a label that differs from the float64 arg-min must stay >= 4x inside tau (headroom), and
with the re-check ON the deferred fraction on ordinary data must stay below 1 % (the bound is not so loose that the
float64 path carries the work).  Formula of tau restated from bkm_api.cu; sizes chosen to run in seconds."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

EPS = 2.0 ** -24


def tau_formula(d, family):
    if family == 3:
        return 2.0 ** -17 + (8.0 * math.sqrt(2.0 * ((d + 7) // 8)) + 16.0) * EPS
    if family == 1:
        return (8.0 * math.sqrt(3.0 * ((d + 7) // 8)) + 16.0) * EPS
    return 8.0 * (math.sqrt(d) + 2.0) * EPS


def make(be, n, d, k, kind, dtype, seed):
    g = torch.Generator(device=be.device).manual_seed(seed)
    if kind == "uniform":
        X = torch.rand((n, d), device=be.device, generator=g) * 20 - 10
    else:
        cent = torch.empty((max(2, k // 2), d), device=be.device).uniform_(-10, 10, generator=g)
        X = cent[torch.randint(0, cent.shape[0], (n,), device=be.device, generator=g)] \
            + torch.randn((n, d), device=be.device, generator=g)
        if kind == "scaled":
            X = X * torch.logspace(-2, 2, d, device=be.device)
    X = be.to_device(X.to(dtype), dtype)          # (n, d) view of a pitch-padded buffer
    C = X[torch.randperm(n, device=be.device, generator=g)[:k]].double().contiguous()
    return X, C


def worst_margin(X, C, labels):
    """largest float64 margin, relative to ||x||^2 + max||c||^2, over rows whose label is not the float64 arg-min"""
    cn = (C * C).sum(1)
    worst, nbad = 0.0, 0
    n = X.shape[0]
    for s in range(0, n, 1 << 17):
        xb = X[s:s + (1 << 17)].double()
        d2 = (xb * xb).sum(1, keepdim=True) + cn[None, :] - 2.0 * xb @ C.T
        best = d2.min(1).values
        got = d2.gather(1, labels[s:s + xb.shape[0]].long()[:, None])[:, 0]
        bad = got > best
        if bool(bad.any()):
            rel = (got[bad] - best[bad]) / ((xb[bad] ** 2).sum(1) + cn.max())
            worst = max(worst, float(rel.max()))
            nbad += int(bad.sum())
    return worst, nbad


CASES = [
    # family, n, d, k, kind, dtype, flags
    (1, 1_000_000, 64, 256, "blobs", torch.float32, "tc"),
    (1, 1_000_000, 64, 256, "uniform", torch.float32, "tc"),
    (1, 1_000_000, 64, 256, "scaled", torch.float32, "tc"),
    (1, 600_000, 41, 100, "blobs", torch.float32, "tc"),
    (2, 1_000_000, 13, 20, "uniform", torch.float32, ""),
    (2, 1_000_000, 16, 31, "blobs", torch.float32, ""),
    (0, 400_000, 24, 40, "uniform", torch.float32, "simt"),
    (3, 400_000, 128, 1024, "blobs", torch.bfloat16, ""),
    (3, 400_000, 64, 300, "uniform", torch.bfloat16, ""),
]


@pytest.mark.parametrize("family,n,d,k,kind,dtype,force", CASES)
def test_tau_headroom_without_recheck(family, n, d, k, kind, dtype, force):
    from dask_ml_b200 import _lib
    from dask_ml_b200.engine import CudaBackend

    base = {"tc": _lib.FLAG_FORCE_TC, "simt": _lib.FLAG_FORCE_SIMT, "": 0}[force]
    be = CudaBackend(flags=base | _lib.FLAG_NO_RECHECK)
    if dtype == torch.bfloat16 and not be.supports_bf16:
        pytest.skip("bf16 path not built")
    if not force:
        assert be.kernel_family(d, k, dtype) == family
    X, C = make(be, n, d, k, kind, dtype, seed=n + d + k)
    pack = be.pack_centers(C, dtype)
    labels = be.empty((n,), torch.int32)
    be.assign_chunk(X, pack, k, labels, None, True, None)
    torch.cuda.synchronize()
    assert int(labels.min()) >= 0 and int(labels.max()) < k
    worst, nbad = worst_margin(X, C, labels)
    tau = tau_formula(d, family)
    assert worst * 4.0 <= tau, "family %d: a label differs from float64 by a relative margin %.3e, tau %.3e" % (family, worst, tau)

    # with the re-check: every label is a float64 arg-min up to genuine float64 near-ties, few rows deferred
    be2 = CudaBackend(flags=base)
    lab2 = be2.empty((n,), torch.int32)
    be2.assign_chunk(X, pack, k, lab2, None, True, None)
    torch.cuda.synchronize()
    worst2, _ = worst_margin(X, C, lab2)
    assert worst2 <= 1e-9
    if family in (1, 3):
        frac = be2.deferred_rows(n, d, k, dtype) / float(n)
        assert frac < 0.01, "deferred fraction %.4f" % frac
