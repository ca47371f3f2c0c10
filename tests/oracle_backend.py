"""TEST-ONLY stand-in for ``dask_ml_b200.engine.CudaBackend`` built on the CPU oracle.

It lets the host-side logic (estimator validation, the Lloyd control flow with the reference's quirks,
k-means|| bookkeeping, the multi-rank all-reduce path over gloo) run in a GPU-less container.  It is never
importable from the product package: tests install it with ``monkeypatch.setattr(k_means, "_BACKEND_FACTORY", ...)``.
"""
import numpy as np
import torch

from oracle import kmeans_oracle as ok


class OracleBackend(object):
    name = "oracle-checker"

    def __init__(self, device=None, flags=0):
        self.device = torch.device("cpu")
        self.flags = flags
        self.launches = 0

    def launch_count(self):
        return self.launches

    def kernel_family(self, d, k, dtype):
        return 0

    def to_device(self, block, dtype):
        t = block if isinstance(block, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(block))
        return t.to(dtype=dtype).contiguous()

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype)

    def zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype)

    def check_finite(self, chunks):
        bad = any((~torch.isfinite(c)).any().item() for c in chunks)
        return torch.tensor([1 if bad else 0], dtype=torch.int32)

    def pack_centers(self, C64, dtype, out=None):
        return C64.clone()

    def _estep(self, x, pack, squared):
        lab, mn = ok.pairwise_distances_argmin_min([x.numpy()], pack.numpy(), metric="euclidean",
                                                   metric_kwargs={"squared": True} if squared else None)
        return lab[0], mn[0]

    def lloyd_chunk(self, x, pack, k, labels, min_d2, sums, counts, inertia):
        self.launches += 1
        lab, mn = self._estep(x, pack, True)
        if labels is not None:
            labels.copy_(torch.from_numpy(lab.astype(np.int32)))
        if min_d2 is not None:
            min_d2.copy_(torch.from_numpy(mn).to(min_d2.dtype))
        s = ok.centers_dense(x.numpy(), lab.astype(np.int32), k)
        sums += torch.from_numpy(s.reshape(-1))
        counts += torch.from_numpy(np.bincount(lab, minlength=k).astype(np.int64))
        if inertia is not None:
            inertia += float(mn.sum())

    def assign_chunk(self, x, pack, k, labels, min_dist, squared, dist_sum):
        self.launches += 1
        lab, mn = self._estep(x, pack, squared)
        if labels is not None:
            labels.copy_(torch.from_numpy(lab.astype(np.int32)))
        if min_dist is not None:
            min_dist.copy_(torch.from_numpy(mn).to(min_dist.dtype))
        if dist_sum is not None:
            dist_sum += float(mn.sum())

    def sample_chunk(self, min_d2, ell_over_phi, seed, row_offset, picked, n_picked):
        n = min_d2.numel()
        u = ok.philox_uniform(int(seed), np.arange(n, dtype=np.uint64) + np.uint64(row_offset))
        hit = np.nonzero(ell_over_phi * min_d2.numpy().astype(np.float64) > u)[0] + row_offset
        base = int(n_picked.item())
        m = min(len(hit), max(0, picked.numel() - base))
        if m:
            picked[base:base + m] = torch.from_numpy(hit[:m].astype(np.int64))
        n_picked += len(hit)

    def min_fold(self, run_min, new_min, phi_acc):
        if new_min is not None:
            torch.minimum(run_min, new_min, out=run_min)
        if phi_acc is not None:
            phi_acc += run_min.sum(dtype=torch.float64)

    def transform_chunk(self, x, pack, k, out, mode=0, gamma=0.0):
        d = ok.euclidean_distances([x.numpy()], pack.numpy().astype(x.numpy().dtype), squared=mode != 0)[0]
        if mode == 2:
            d = np.exp(-gamma * d)
        out.copy_(torch.from_numpy(d).to(out.dtype))

    def finalize(self, sums, counts, C_old, C_new, shift):
        k, d = C_old.shape
        c = torch.clamp(counts, min=1).to(torch.float64)
        C_new.copy_(sums.view(k, d) / c[:, None])
        shift.copy_(((C_old - C_new) ** 2).sum().reshape(1))
