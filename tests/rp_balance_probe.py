"""Manual probe: how unevenly do the label buckets of the M-step row pass fill?  (not collected by pytest)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import synth_config_device
from dask_ml_b200.cluster.k_means import LloydState
from dask_ml_b200.engine import Comm, CudaBackend, DeviceData
be = CudaBackend()
X = synth_config_device("C5s", 4_000_000, 0, be.device)
data = DeviceData([X], be, Comm())
k = 1024
st = LloydState(data, data.global_rows(list(range(k))).astype(np.float64))
for it in range(12):
    st.step(); st.accept()
    if it in (0, 3, 11):
        lab = st.labels[0].long()
        cnt = torch.bincount(lab, minlength=k).double()
        CS = 4
        cs = lab & (CS - 1); cl = lab >> 2; w = cl & 15
        b = (cs * 16 + w)
        bc = torch.bincount(b, minlength=64).double().view(4, 16)
        print("iter", it, "cluster size mean %.0f max %.0f cv %.2f empty %d" % (cnt.mean(), cnt.max(), cnt.std() / cnt.mean(), int((cnt == 0).sum())),
              "| per-slice warp load max/mean:", [round(float(bc[s].max() / bc[s].mean()), 2) for s in range(4)],
              "| slice totals", [int(v) for v in bc.sum(1)])
        # per row block of 1/37 of the rows
        nb = 37; per = (lab.numel() + nb - 1) // nb
        worst = []
        for r in range(0, nb, 9):
            bb = torch.bincount(b[r * per:(r + 1) * per], minlength=64).double().view(4, 16)
            worst.append(round(float((bb.max(1).values / bb.mean(1)).max()), 2))
        print("    row-block max/mean (sampled blocks):", worst)
