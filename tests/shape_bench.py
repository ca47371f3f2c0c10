"""Side measurements (not collected by pytest, not the bench contract): Lloyd-iteration time of the other
BASELINE.json shapes on one GPU, against their HBM roofline (algorithmic bytes d*s+4 per sample).
    python tests/shape_bench.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bench import synth_blobs_device
from dask_ml_b200.cluster.k_means import LloydState
from dask_ml_b200.engine import Comm, CudaBackend, DeviceData

peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
be = CudaBackend()
out = []
for name, n, d, k, dt in [("C3 KDD-shaped 4.9M x 41 f32 k=100", 4_898_431, 41, 100, torch.float32),
                          ("C4 airline-shaped shard 15M x 13 f32 k=20 (1/8 of 120M)", 15_000_000, 13, 20, torch.float32),
                          ("C1 100k x 16 f64 k=8", 100_000, 16, 8, torch.float64),
                          ("C2 on the CUDA-core kernel (FORCE_SIMT) 2M x 64 f32 k=256", 2_000_000, 64, 256, torch.float32)]:
    be.flags = 1 if name.startswith("C2") else 0
    X = be.to_device(synth_blobs_device(n, d, k, 7, be.device, dt), dt)      # padded row pitch when d % 4 != 0
    data = DeviceData([X], be, Comm())
    st = LloydState(data, X[:k].cpu().numpy().astype(np.float64))
    for _ in range(3):
        st.step(); st.accept()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    steps = 10
    e0.record()
    for _ in range(steps):
        st.step(); st.accept()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    s = 4 if dt == torch.float32 else 8
    gbs = n * (d * s + 4) / (ms * 1e-3) / 1e9
    tf = 2.0 * n * d * k / (ms * 1e-3) / 1e12
    fam = int(be.kernel_family(d, k, dt)) if X.stride(0) % 4 == 0 else 0
    rec = {"shape": name, "kernel_family": fam, "row_pitch": int(X.stride(0)), "ms_per_iter": ms,
           "samples_per_s": n / (ms * 1e-3), "hbm_gbs": gbs, "hbm_frac_of_measured": gbs / peaks["hbm_gbs"], "tflops": tf}
    print(json.dumps(rec), flush=True)
    del X, data, st
    torch.cuda.empty_cache()
