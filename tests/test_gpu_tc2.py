"""GPU parity tests of the large-shape tensor path (bkm_tc2.cu + bkm_rowpass.cu): bfloat16 rows, any k, d <= 128
(BASELINE config C5: 128 features, k = 1024).  The oracle for bf16 rows is the float64 E-step on the SAME values (every
bf16 number is exactly representable in float32 / float64)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    from dask_ml_b200.engine import CudaBackend

    return CudaBackend()


def _exact(X64, C):
    import torch

    cn = (C * C).sum(1)
    out = torch.empty((X64.shape[0],), dtype=torch.int64, device=X64.device)
    margin = torch.empty((X64.shape[0],), dtype=torch.float64, device=X64.device)
    for s in range(0, X64.shape[0], 1 << 16):
        xb = X64[s:s + (1 << 16)]
        d2 = (xb * xb).sum(1, keepdim=True) + cn[None, :] - 2.0 * xb @ C.T
        if C.shape[0] > 1:
            top = torch.topk(d2, 2, dim=1, largest=False)
            out[s:s + xb.shape[0]] = top.indices[:, 0]
            margin[s:s + xb.shape[0]] = top.values[:, 1] - top.values[:, 0]
        else:
            out[s:s + xb.shape[0]] = 0
            margin[s:s + xb.shape[0]] = 1.0
    return out, margin


def _data(be, n, d, k, seed, spread=10.0):
    import torch

    g = torch.Generator(device=be.device).manual_seed(seed)
    cent = torch.empty((max(2, k // 2), d), device=be.device).uniform_(-spread, spread, generator=g)
    X = (cent[torch.randint(0, cent.shape[0], (n,), device=be.device, generator=g)] +
         torch.randn((n, d), device=be.device, generator=g)).to(torch.bfloat16)
    x = be.to_device(X, torch.bfloat16)
    C = X[torch.randint(0, n, (k,), device=be.device, generator=g)].double() + \
        0.01 * torch.randn((k, d), device=be.device, generator=g, dtype=torch.float64)
    return X, x, C.contiguous()


SHAPES = [(5000, 128, 1024), (3001, 64, 300), (1000, 128, 16), (129, 16, 1), (70001, 128, 1000), (20000, 96, 512),
          (4000, 100, 257), (1, 8, 3), (127, 128, 256), (148 * 128 * 2 + 5, 32, 40), (300_000, 128, 1024)]


@pytest.mark.parametrize("n,d,k", SHAPES)
def test_tc2_lloyd_chunk_bf16(be, n, d, k):
    import torch

    assert be.kernel_family(d, k, torch.bfloat16) == 3
    X, x, C = _data(be, n, d, k, n + d + k)
    assert x.stride(0) % 8 == 0
    pack = be.pack_centers(C, torch.bfloat16)
    labels = be.empty((n,), torch.int32)
    mind2 = be.empty((n,), torch.float32)
    sums = be.zeros((k * d,), torch.float64)
    counts = be.zeros((k,), torch.int64)
    inertia = be.zeros((1,), torch.float64)
    for _ in range(2):          # twice: a warm SM reuses every barrier / ring
        sums.zero_(); counts.zero_(); inertia.zero_()
        be.lloyd_chunk(x, pack, k, labels, mind2, sums, counts, inertia)
    torch.cuda.synchronize()
    X64 = X.double()
    want, margin = _exact(X64, C)
    got = labels.long()
    assert int(got.min()) >= 0 and int(got.max()) < k
    bad = got != want
    if bool(bad.any()):
        xs = (X64 ** 2).sum(1)[bad] + (C ** 2).sum(1).max()
        assert bool((margin[bad] <= 1e-9 * xs).all()), (int(bad.sum()), float((margin[bad] / xs).max()))
    assert torch.equal(counts, torch.bincount(got, minlength=k))
    ref = torch.zeros((k, d), dtype=torch.float64, device=be.device).index_add_(0, got, X64)
    assert float((sums.view(k, d) - ref).abs().max()) <= 2e-6 * float(ref.abs().max()) + 1e-9
    exact = ((X64 - C[got]) ** 2).sum(1)
    scale = (X64 ** 2).sum(1) + (C ** 2).sum(1).max()
    assert float(((mind2.double() - exact).abs() / scale).max()) < 2e-6
    assert abs(float(inertia[0]) - float(exact.sum())) <= 1e-5 * float(exact.sum()) + 2e-6 * float(scale.sum())
    # assignment only (labels only, then labels + non-squared distances); Lloyd step without labels / distances
    lab2 = be.empty((n,), torch.int32)
    be.assign_chunk(x, pack, k, lab2, None, True, None)
    md = be.empty((n,), torch.float32)
    ds = be.zeros((1,), torch.float64)
    lab3 = be.empty((n,), torch.int32)
    be.assign_chunk(x, pack, k, lab3, md, False, ds)
    sums2 = be.zeros((k * d,), torch.float64)
    counts2 = be.zeros((k,), torch.int64)
    be.lloyd_chunk(x, pack, k, None, None, sums2, counts2, None)
    torch.cuda.synchronize()
    assert torch.equal(lab2, labels) and torch.equal(lab3, labels)
    assert float(((md.double() ** 2 - exact).abs() / scale).max()) < 4e-6
    assert torch.equal(counts2, counts) and torch.equal(sums2, sums)
    # the rounding bound defers only a small fraction of rows (these centres are rows of X, two per blob on average:
    # the blobs they split are full of genuine near-ties, a few percent of the rows sit inside the 16-bit bound)
    if n >= 5000:
        assert be.deferred_rows(n, d, k, torch.bfloat16) <= 0.08 * n


def test_tc2_ties_and_extremes(be):
    """Duplicate centres -> lowest index; rows with huge / non-finite entries take the float64 path."""
    import torch

    n, d, k = 4096, 128, 600
    X, x, C = _data(be, n, d, k, 5)
    C[300] = C[7]
    C[599] = C[0]
    Xm = X.clone()
    Xm[11] = 1.0e30
    Xm[13, 5] = float("inf")
    x = be.to_device(Xm, torch.bfloat16)
    pack = be.pack_centers(C, torch.bfloat16)
    labels = be.empty((n,), torch.int32)
    be.assign_chunk(x, pack, k, labels, None, True, None)
    torch.cuda.synchronize()
    got = labels.long()
    assert not bool(torch.isin(got, torch.tensor([300, 599], device=be.device)).any())
    want, margin = _exact(Xm.double(), C)
    bad = got != want
    bad[11] = False; bad[13] = False            # no float64 answer either (inf - inf)
    assert int(bad.sum()) == 0 or bool((margin[bad] <= 1e-9 * ((Xm.double() ** 2).sum(1)[bad] + (C ** 2).sum(1).max())).all())


def test_tc2_fit_matches_oracle(be, oracle):
    """KMeans.fit on bf16 rows against the CPU oracle on the same values (float32 view), identical init."""
    import torch
    from dask_ml_b200.cluster import KMeans

    n, d, k = 30000, 128, 300
    X, x, _ = _data(be, n, d, k, 77)
    Xh = X.float().cpu().numpy()
    init = Xh[:k].copy()
    km = KMeans(n_clusters=k, init=init, max_iter=6, tol=0.0).fit(X)
    lab, inertia, C, n_iter = oracle.kmeans_single_lloyd(oracle.to_blocks(Xh, 10000), k, init=init, max_iter=6, tol=0.0)
    got = km.labels_.compute()
    want = np.concatenate(lab)
    assert km.n_iter_ == n_iter
    assert int((got != want).sum()) <= 3
    assert abs(km.inertia_ - inertia) / inertia < 1e-4
    assert km.cluster_centers_.dtype == np.float32
    np.testing.assert_allclose(km.cluster_centers_, C, rtol=1e-4, atol=1e-4)
    pred = km.predict(X).compute()
    assert int((pred != got).sum()) <= 3


def test_tc2_fit_chunked_equals_single_chunk(be, oracle):
    """The C5 layout of bench.py: several resident bf16 chunks per GPU (ragged sizes).  Labels / inertia / centres must
    not depend on how the rows are chunked (up to the fp32 partial sums), and predict / transform follow."""
    import torch
    from dask_ml_b200.chunked import ChunkedArray
    from dask_ml_b200.cluster import KMeans

    n, d, k = 41000, 128, 600
    X, x, _ = _data(be, n, d, k, 91)
    init = X[:k].float().cpu().numpy()
    one = KMeans(n_clusters=k, init=init, max_iter=4, tol=0.0).fit(X)
    cuts = [0, 9000, 9001, 25000, n]
    parts = ChunkedArray([X[a:b] for a, b in zip(cuts, cuts[1:])])
    many = KMeans(n_clusters=k, init=init, max_iter=4, tol=0.0).fit(parts)
    assert many.n_iter_ == one.n_iter_
    la, lb = one.labels_.compute(), many.labels_.compute()
    assert int((la != lb).sum()) <= 3
    assert abs(many.inertia_ - one.inertia_) / one.inertia_ < 1e-5
    np.testing.assert_allclose(many.cluster_centers_, one.cluster_centers_, rtol=2e-3, atol=2e-3)
    # predict = labels of the fitted centres on the same rows (Q4: shift > 1e-7 -> labels_ come from a re-label)
    pred = many.predict(parts).compute()
    assert int((pred != lb).sum()) <= 3
    # against the float64 arg-min of the final centres
    C = torch.as_tensor(many.cluster_centers_.astype(np.float64)).to(be.device)
    want, margin = _exact(X.double(), C)
    bad = torch.as_tensor(pred).to(be.device).long() != want
    assert int(bad.sum()) == 0 or bool((margin[bad] <= 1e-9 * ((X.double() ** 2).sum(1)[bad] + (C ** 2).sum(1).max())).all())
