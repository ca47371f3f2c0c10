"""Manual probe (not collected by pytest): a few Lloyd chunk calls of the large-shape tensor path, for ncu.
    python tests/tc2_probe.py [n] [d] [k]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dask_ml_b200.engine import CudaBackend
from bench import synth_blobs_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
k = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
be = CudaBackend()
X = synth_blobs_device(n, d, k, 3, be.device, torch.bfloat16)
C = X[:k].double().contiguous()
pack = be.pack_centers(C, torch.bfloat16)
labels = be.empty((n,), torch.int32)
sums = be.zeros((k * d,), torch.float64); counts = be.zeros((k,), torch.int64)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for it in range(4):
    sums.zero_(); counts.zero_()
    if it == 1: ev[0].record()
    be.lloyd_chunk(X, pack, k, labels, None, sums, counts, None)
ev[1].record(); torch.cuda.synchronize()
print("ms per Lloyd chunk call: %.3f  (n=%d d=%d k=%d) deferred=%s" % (ev[0].elapsed_time(ev[1]) / 3, n, d, k, be.deferred_rows(n, d, k, torch.bfloat16)))
